/*
 * pbd_b200.h -- C ABI of libpbd_b200.so: the B200 (sm_100a) replacement for the PBDServer
 * XPBD substep of Captain-Noble/CS121-softbodysim.
 *
 * This is the drop-in boundary.  Plain C types only, caller owns every host
 * pointer, the library copies what it needs at create time; integer status codes (0 = ok);
 * no exceptions cross the boundary.  A handle is single-threaded (the reference calls
 * step()/pack_positions() from exactly one thread, CProgram/src/Sim.cpp:386-390).  All solver
 * state (x, v, xStar, w, rest values, lambdas, the colour/tile schedule) lives in HBM between
 * calls.  There is NO CPU fallback: every entry that computes needs a CUDA device.
 *
 * What each entry replaces in the reference (paths relative to /root/reference):
 *
 *   pbd_create           the MSG_INIT decode + init helpers:  CProgram/src/Server.cpp:30-114
 *   pbd_create_from_init the same, straight from the wire payload:  CProgram/src/Server.cpp:30-70
 *                        (builds PBDState), CProgram/src/Sim.cpp:63-79 compute_inv_mass and
 *                        :81-95 build_rest.  Inverse masses and rest values are derived on the
 *                        host in the CALLER'S constraint order, exactly as the reference does.
 *   pbd_step             IStepper::step           CProgram/include/PBDServer.h:264,
 *                        SerialStepper::step      CProgram/src/Sim.cpp:280-305
 *                        (predict :178-185, solve_edges_xpbd_gs :100-130, solve_tets_xpbd_gs
 *                        :132-173, project_ground :187-195, commit :197-222).
 *   pbd_read_positions   IStepper::pack_positions CProgram/include/PBDServer.h:265,
 *                        CProgram/src/Sim.cpp:307-316: committed x of ALL V vertices in the
 *                        caller's vertex order, 3V floats (the MSG_POSITIONS payload,
 *                        CProgram/src/Server.cpp:10-18).
 *   pbd_backend_name     IStepper::name           CProgram/include/PBDServer.h:263.
 *   pbd_destroy          PBDState going out of scope / a second MSG_INIT replacing it
 *                        (CProgram/src/Server.cpp:106-110).
 *   pbd_params           struct SolverParams      CProgram/include/PBDServer.h:147-161, in the
 *                        MSG_INIT wire order      CProgram/src/Server.cpp:38-50.
 *   pbd_step_stats       perf::StepStats          CProgram/include/PBDServer.h:75-81.
 *   pbd_batch_*          the same three calls for many independent bodies (BASELINE.json
 *                        config 4); the reference runs one body per process.
 *
 * The order in which constraints are projected is a conflict-free parallel schedule, i.e. a
 * permutation of the reference's array order (Gauss-Seidel over the same constraint set; the
 * permutation is disclosed by pbd_get_schedule_order so the reference itself can be run on the
 * permuted arrays and compared bit for bit -- tests/test_parity_gpu.py).
 */
#ifndef PBD_B200_H
#define PBD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBD_ABI_VERSION 1

/* status codes */
enum {
  PBD_OK = 0,
  PBD_ERR_INVALID = 1,    /* null pointer, bad size, bad option                      */
  PBD_ERR_INDEX = 2,      /* an edge/tet index >= V (the reference would read out of bounds) */
  PBD_ERR_NO_DEVICE = 3,  /* no usable CUDA device: there is no CPU fallback         */
  PBD_ERR_CUDA = 4,       /* a CUDA runtime call failed; see pbd_last_error()         */
  PBD_ERR_OOM = 5,
  PBD_ERR_UNSUPPORTED = 6
};

/* SolverParams, MSG_INIT wire order (48 bytes, no padding). omega and dtHint are carried but,
 * as in the reference, never read by the solver. */
typedef struct pbd_params {
  uint32_t substeps;          /* clamped to >= 1 at step time (Sim.cpp:285)     */
  uint32_t iterations;        /* may be 0 (Sim.cpp:293)                          */
  float dtHint;
  float omega;
  float edgeCompliance;       /* alpha = max(0,c)/dt^2                           */
  float volumeCompliance;
  float gx, gy, gz;
  uint32_t groundEnabled;
  float groundY;
  float friction;             /* clamped to [0,1] at commit                      */
} pbd_params;

/* perf::StepStats. pbd_step ADDS into these (the caller zero-initialises per frame, as
 * sim_thread_fn does, Sim.cpp:386).  The tile and batch backends run the whole frame as ONE kernel with
 * predict / ground / commit fused into vertex stages of it: predictMs and commitMs are the frame's
 * device time x the share of CTA cycles spent in the vertex stages that carry that work (a fused
 * commit+predict stage between two substeps is charged half to each; the waits of the hand-over that
 * fall inside such a stage are part of it), solveMs is the rest, so the three add up to the frame's
 * device time.  The stream backend times its stage kernels with events under PBD_FLAG_STAGE_TIMING
 * (otherwise everything is in solveMs).  totalMs is host wall time of the call. */
typedef struct pbd_step_stats {
  double predictMs, solveMs, commitMs, packMs, totalMs;
} pbd_step_stats;

enum { PBD_BACKEND_AUTO = 0, PBD_BACKEND_STREAM = 1, PBD_BACKEND_TILE = 2,
       PBD_BACKEND_JACOBI = 3 /* comparison schedule, NOT the PBDServer algorithm: the Jacobi + SOR gather solver of the
                                 reference's in-engine path (SoftBodySolver.cs:379-527, SoftBodyCompute.compute:229-389):
                                 two grid-wide phases per constraint type and iteration, stiffness in [0,1] instead of
                                 XPBD compliance, SOR factor = pbd_params.omega (the reference default there is 1.4) */ };
enum { PBD_ORDER_STRICT = 0, PBD_ORDER_INTERLEAVED = 1, PBD_ORDER_RIDING = 2 };
enum {
  PBD_FLAG_STAGE_TIMING = 1u << 0, /* stream backend: no CUDA graph, CUDA events per stage   */
  PBD_FLAG_NO_GRAPH = 1u << 1,     /* stream backend: plain launches (debug / compute-sanitizer) */
  PBD_FLAG_TAGGED_HANDOVER = 1u << 2, /* tile backend, EXPERIMENTAL: positions travel between
                                      tiles as 64-bit {value, tag} pairs, no release fence / done flags.
                                      Used only when every phase covers every vertex and the body is on
                                      one GPU; otherwise ignored.                                          */
  PBD_FLAG_FAST_ARITH = 1u << 3    /* tile and batch backends: the projections use FFMA, folded 1/6 factors and
                                      the SFU reciprocal / rsqrt instead of the reference's SSE2 rounding
                                      sequence (csrc/pbd_math.cuh: *_delta_fast).  Same constraints, same
                                      schedule, results equal to the exact mode up to rounding: validated by
                                      tolerance (RMS <= 1e-4 of the bounding-box diagonal after 10 frames,
                                      residuals after 1000 frames), not bit for bit.                          */
};

typedef struct pbd_options {
  uint32_t struct_size;    /* = sizeof(pbd_options); lets the struct grow                       */
  uint32_t backend;        /* PBD_BACKEND_*                                                      */
  uint32_t order_mode;     /* PBD_ORDER_STRICT: every iteration projects all edges, then all tets,
                              then the ground clamp -- the reference's sweep order up to a
                              permutation inside each constraint type.
                              PBD_ORDER_INTERLEAVED (tile backend): every iteration still projects
                              every constraint exactly once, but edges and tets interleave -- a
                              tile visit runs colour steps that each hold a vertex-disjoint set of
                              its edges and tets.  pbd_get_schedule_sequence discloses the order.
                              PBD_ORDER_RIDING (tile backend): interleaved, and an edge whose two vertices
                              belong to a tet of the same tile visit RIDES on that tet -- the tet's thread
                              projects it right after the tet (at most two riders per tet, opposite edges),
                              so it needs no colour step and no vertex gather of its own.  Still one
                              projection of every constraint per iteration, i.e. a permutation of the
                              reference's order; disclosed by pbd_get_schedule_sequence like the others. */
  uint32_t flags;          /* PBD_FLAG_*                                                         */
  uint32_t tile_vertices;  /* tile backend: target vertices per shared-memory tile, 0 = auto     */
  uint32_t block_threads;  /* 0 = auto                                                           */
  uint32_t max_phases;     /* reserved (ignored)                                                 */
  uint32_t partitions;     /* tile backend: shifted vertex partitions per sweep, 0 = auto (4)    */
  uint32_t lanes_per_tet;  /* tile backend: 1, 2 or 4 lanes cooperate on one tet, 0 = auto (1)   */
  uint32_t tiles_per_sm;   /* tile backend: tiles (CTAs) resident per SM, 0 = auto               */
  uint32_t shard_world;    /* one body across several GPUs of a node: number of ranks (0/1 = off) */
  uint32_t shard_rank;     /* ... and which of them this handle is                               */
  uint32_t plan_sms;       /* plan for this many SMs, 0 = the device's SM count x shard_world     */
  float jacobi_edge_stiffness;    /* PBD_BACKEND_JACOBI: edgeStiffness (0 = the reference default 0.9)    */
  float jacobi_volume_stiffness;  /* PBD_BACKEND_JACOBI: volumeStiffness (0 = the reference default 0.98) */
  uint32_t reserved[1];
} pbd_options;

typedef struct pbd_info {
  uint32_t V, E, T;
  uint32_t backend;             /* resolved PBD_BACKEND_*                                    */
  uint32_t edge_colors;         /* stream: global colours; tile: max local colours summed over phases */
  uint32_t tet_colors;
  uint32_t edge_phases;         /* tile backend: grid-wide phases per iteration that carry edges */
  uint32_t tet_phases;
  uint32_t tiles;               /* tile backend: total tiles over all phases                 */
  uint32_t launches_per_frame;  /* kernels launched by one pbd_step at the current params    */
  uint32_t grid_blocks, block_threads;
  uint32_t partitions;          /* tile backend: shifted vertex partitions (main phases per sweep) */
  uint32_t lanes_per_tet;       /* tile backend: lanes cooperating on one tet                      */
  uint32_t gather_wavefronts_permille[2]; /* tile backend, one thread per constraint: 1000 x shared-memory wavefronts per
                                             quarter-warp vertex gather of the sweeps under this schedule's placement
                                             (edges, tets; 1000 = free of bank conflicts; csrc/pbd_placement.cpp); 0 = n/a */
  uint64_t device_bytes;        /* HBM held by this handle                                    */
  uint64_t algorithmic_bytes_per_substep; /* 104 V + I (20 E + 28 T + 84 V), SURVEY.md 8(d)   */
  double plan_ms;               /* host time spent building the schedule                     */
  double upload_ms;
} pbd_info;

typedef struct pbd_handle pbd_handle;
typedef struct pbd_plan pbd_plan;
typedef struct pbd_batch pbd_batch;

int pbd_abi_version(void);
/* thread-local message of the last failing call */
const char* pbd_last_error(void);
/* number of CUDA devices visible, or 0 (never fails) */
int pbd_device_count(void);

/* ---- single body -------------------------------------------------------------------- */

/* x0: 3V world-space floats; edgeIds: 2E; tetIds: 4T; pinned: nPinned vertex indices (entries
 * >= V are ignored, Sim.cpp:76-78).  opts may be NULL.  device < 0 = current device.
 * Returns NULL and sets *status (if non-NULL) on failure. */
pbd_handle* pbd_create(const pbd_params* params, uint32_t V, uint32_t E, uint32_t T,
                       const float* x0, const uint32_t* edgeIds, const uint32_t* tetIds,
                       const uint32_t* pinned, uint32_t nPinned, int device,
                       const pbd_options* opts, int* status);

/* The same from the raw MSG_INIT payload of the reference's wire protocol "PBD1" -- the decode that
 * comm_loop performs (CProgram/src/Server.cpp:30-70; written by PBDRemoteWorld.cs:294-349).  Little-
 * endian, packed: u32 V, E, T | pbd_params (48 bytes) | u32 pinnedCount | u32 pinned[pinnedCount]
 * | f32 x0[3V] | u32 edgeIds[2E] | u32 tetIds[4T].  The payload need not be aligned.  Unlike the
 * reference, which trusts V/E/T, a payload shorter than its own counts demand is refused with
 * PBD_ERR_INVALID instead of being read past its end; trailing bytes are ignored as there. */
pbd_handle* pbd_create_from_init(const void* payload, uint64_t size, int device, const pbd_options* opts,
                                 int* status);
/* bytes of a MSG_INIT payload with these counts (what the client computes, PBDRemoteWorld.cs:294-306) */
uint64_t pbd_init_payload_size(uint32_t V, uint32_t E, uint32_t T, uint32_t pinnedCount);

/* Advance one frame: max(1,substeps) substeps of dt/substeps.  Synchronous (returns after the
 * device finished), like IStepper::step.  stats may be NULL. */
int pbd_step(pbd_handle* h, float dt, pbd_step_stats* stats);
/* Enqueue `frames` frames without synchronising; pbd_sync waits and reports device ms. */
int pbd_step_async(pbd_handle* h, float dt, uint32_t frames);
int pbd_sync(pbd_handle* h, double* device_ms);

/* 3V floats, caller's vertex order, committed positions x (not xStar). Host pointer. */
int pbd_read_positions(pbd_handle* h, float* out, double* packMs);

void pbd_destroy(pbd_handle* h);
const char* pbd_backend_name(const pbd_handle* h);
int pbd_get_info(const pbd_handle* h, pbd_info* out);
int pbd_set_params(pbd_handle* h, const pbd_params* params); /* substeps/iterations/compliances/gravity/ground */

/* Per-vertex normals on the GPU, so that a client need not RecalculateNormals on the CPU
 * (PBDRemoteSoftBody.cs:203-211): K_UpdateNormals of Assets/Shaders/SoftBodyCompute.compute:459-491 -- the
 * normalised sum of cross(pb - pa, pc - pa) over the surface triangles incident to a vertex, in
 * BuildTriAdjacency's order (SoftBodySolver.cs:1173-1213); (0, 1, 0) for vertices on no triangle.
 * pbd_set_surface: nTris triangles (3 vertex indices each, caller's numbering; the asset's surfaceTriIds).
 * pbd_read_normals: 3V floats, caller's vertex order, from the committed positions. */
int pbd_set_surface(pbd_handle* h, const uint32_t* surfaceTriIds, uint32_t nTris);
int pbd_read_normals(pbd_handle* h, float* out);

/* Primitive colliders in the clamp stage (BASELINE.json north_star kernel (3) "ground and
 * collision clamping").  PBDServer has only the y-plane (Sim.cpp:187-195); the sphere / oriented box / capsule
 * push-out is the reference's in-engine solver's: Assets/Scripts/Softbody/SoftBodyCollisionMath.cs:8-110,
 * collider data as SoftBodyPrimitiveCollider.PrimitiveColliderData (SoftBodyPrimitiveCollider.cs:8-14), applied
 * like SoftBodySolver.cs:554-561 -- after the ground clamp of every iteration, colliders in order, on every vertex
 * with mass: if it penetrates (within particleRadius), it is moved out along the minimal translation.
 * n = 0 removes them.  Single bodies only (tile and stream backends). */
enum { PBD_COLLIDER_SPHERE = 0, PBD_COLLIDER_BOX = 1, PBD_COLLIDER_CAPSULE = 2 };
#define PBD_MAX_COLLIDERS 16
typedef struct pbd_collider {
  uint32_t type;          /* PBD_COLLIDER_*                                                    */
  float px, py, pz;       /* positionW                                                         */
  float qx, qy, qz, qw;   /* rotationW, unit quaternion (capsule axis = its local Y)           */
  float dx, dy, dz;       /* sphere: radius | box: half extents | capsule: radius, half height */
} pbd_collider;
int pbd_set_colliders(pbd_handle* h, const pbd_collider* colliders, uint32_t n, float particleRadius);

/* The projection order actually executed, as indices into the caller's arrays:
 * edgeOrder[k] = caller edge projected k-th inside an iteration's edge sweep (E entries),
 * tetOrder likewise (T entries).  Either pointer may be NULL. */
int pbd_get_schedule_order(const pbd_handle* h, uint32_t* edgeOrder, uint32_t* tetOrder);
/* interleaved order mode: the full per-iteration sequence, E+T entries, bit 31 set = tet,
 * low bits = position in edgeOrder / tetOrder. */
int pbd_get_schedule_sequence(const pbd_handle* h, uint32_t* items);

/* parity / debug readback in the CALLER'S indexing */
enum {
  PBD_ARRAY_INV_MASS = 0,    /* V floats  */
  PBD_ARRAY_EDGE_REST = 1,   /* E floats  */
  PBD_ARRAY_TET_REST = 2,    /* T floats  */
  PBD_ARRAY_EDGE_LAMBDA = 3, /* E floats  */
  PBD_ARRAY_TET_LAMBDA = 4,  /* T floats  */
  PBD_ARRAY_VELOCITY = 5,    /* 3V floats */
  PBD_ARRAY_XSTAR = 6        /* 3V floats */
};
int pbd_get_array(pbd_handle* h, int what, float* out);

/* ---- one body across several GPUs of one node (BASELINE.json config 5) -------------------
 *
 * Every rank (one process per GPU, or one handle per device in a single process) calls pbd_create
 * with the SAME mesh and opts.shard_world / opts.shard_rank set.  All ranks build the same
 * schedule; a rank steps the tiles it owns.  Vertices live on the rank that owns their home tile
 * and are read and written IN PLACE by the tiles of other ranks over NVLink (peer pointers); the
 * per-tile done counters that order the tiles work across GPUs the same way (system-scope
 * release/acquire).  There is no separate halo-exchange step and the result is bit-identical to
 * a single-GPU run of the same schedule (opts.plan_sms).
 *
 *   multi-process:  pbd_shard_export -> all-gather the 128-byte blobs -> pbd_shard_attach_ipc
 *   one process:    pbd_shard_attach_local(handles, world)
 * then pbd_step_async on every rank (all ranks must be launched before any is synchronised),
 * pbd_sync, and pbd_read_positions, which returns valid positions for the vertices this rank
 * owns (pbd_shard_owner tells which). */
#define PBD_SHARD_EXPORT_BYTES 128
int pbd_shard_export(pbd_handle* h, void* out /* PBD_SHARD_EXPORT_BYTES */);
int pbd_shard_attach_ipc(pbd_handle* h, const void* all /* world x PBD_SHARD_EXPORT_BYTES, rank order */);
int pbd_shard_attach_local(pbd_handle* const* handles, uint32_t world);
int pbd_shard_owner(const pbd_handle* h, uint8_t* ownerOfVertex /* V */);

/* ---- schedule only (pure host code, no CUDA: usable on a machine without a GPU) ------- */

pbd_plan* pbd_plan_create(uint32_t V, uint32_t E, uint32_t T, const float* x0,
                          const uint32_t* edgeIds, const uint32_t* tetIds,
                          const pbd_options* opts, int* status);
int pbd_plan_get_info(const pbd_plan* p, pbd_info* out);
int pbd_plan_get_order(const pbd_plan* p, uint32_t* edgeOrder, uint32_t* tetOrder);
int pbd_plan_get_sequence(const pbd_plan* p, uint32_t* items);
/* per-constraint (phase, tile, colour) of the schedule, caller indexing; any pointer may be NULL.
 * Stream backend: phase = 0, tile = 0, colour = global colour. */
int pbd_plan_get_edge_slots(const pbd_plan* p, uint32_t* phase, uint32_t* tile, uint32_t* colour);
int pbd_plan_get_tet_slots(const pbd_plan* p, uint32_t* phase, uint32_t* tile, uint32_t* colour);
void pbd_plan_destroy(pbd_plan* p);

/* ---- batch of independent bodies (BASELINE.json config 4) ----------------------------- */

/* Bodies are concatenated: body b owns vertices [vOff[b], vOff[b+1]), edges [eOff[b], eOff[b+1]),
 * tets [tOff[b], tOff[b+1]); edge/tet indices are LOCAL to the body.  All bodies share params.
 * One kernel per frame, each CTA steps whole bodies out of shared memory (csrc/pbd_batch.cu), so a
 * body must fit one SM (16 B/vertex + 20 B/edge + 28 B/tet <= ~220 KB, < 65536 vertices);
 * otherwise PBD_ERR_UNSUPPORTED -- use pbd_create for such a body. */
pbd_batch* pbd_batch_create(const pbd_params* params, uint32_t nBodies, const uint64_t* vOff,
                            const uint64_t* eOff, const uint64_t* tOff, const float* x0,
                            const uint32_t* edgeIds, const uint32_t* tetIds, int device,
                            const pbd_options* opts, int* status);
int pbd_batch_step(pbd_batch* b, float dt, pbd_step_stats* stats);
int pbd_batch_step_async(pbd_batch* b, float dt, uint32_t frames);
int pbd_batch_sync(pbd_batch* b, double* device_ms);
int pbd_batch_read_positions(pbd_batch* b, float* out, double* packMs); /* 3*vOff[nBodies] floats */
int pbd_batch_get_info(const pbd_batch* b, pbd_info* out);
/* per body: edges colour by colour, then tets colour by colour; entries are BODY-LOCAL constraint
 * indices, concatenated in body order (eOff / tOff positions).  Either pointer may be NULL. */
int pbd_batch_get_schedule_order(const pbd_batch* b, uint32_t* edgeOrder, uint32_t* tetOrder);
void pbd_batch_destroy(pbd_batch* b);

#ifdef __cplusplus
}
#endif
#endif /* PBD_B200_H */
