// pbd_server.cpp -- the "PBD1" wire server around the GPU stepper.
//
// A from-scratch Linux implementation of the reference's server process
// (CProgram/src/main.cpp:69-98, Server.cpp:20-149, Net.cpp:57-102) on top of the C ABI
// (include/pbd_b200.h), so that the stock Unity client (Assets/Scripts/Softbody/PBDRemoteWorld.cs)
// talks to the B200 solver unchanged:
//
//   header  both   u32 magic 'PBD1' = 0x31444250 | u32 type | u32 payload bytes      PBDServer.h:47-62
//   INIT=1  C->S   64 fixed bytes + pinned + x0 + edgeIds + tetIds, no reply         Server.cpp:30-114
//   STEP=2  C->S   f32 dt (size >= 4, else disconnect); before INIT: ignored, no reply  Server.cpp:115-136
//   POSITIONS=3 S->C  f32 pos[3V]: committed x of ALL vertices, caller order         Server.cpp:10-18, Sim.cpp:307-316
//   SHUTDOWN=4  C->S  server exits; unknown type: disconnect                         Server.cpp:138-143
//
// What differs from the reference, on purpose:
//   * the MSG_INIT payload is validated (pbd_create_from_init: size against V/E/T/pinnedCount, indices
//     < V, finite positions); a bad INIT is answered by closing the connection -- the reference trusts
//     it and reads out of bounds;
//   * TCP_NODELAY on the accepted socket and ONE gathered send (header + payload in a single sendmsg)
//     per MSG_POSITIONS: the reference's two send() calls without NODELAY stall ~40 ms per frame on
//     Linux loopback (Nagle + delayed ACK);
//   * one thread: the protocol has strictly one STEP in flight (PBDRemoteWorld.cs:201-246), so the
//     reference's sim thread + condition variables (Sim.cpp:366-398) have nothing to overlap with;
//   * the 1 Hz stats line (Sim.cpp:412-417) is kept field for field and extended with substeps/s,
//     tet-constraints/s and algorithmic GB/s.  predict / commit are vertex stages of the ONE frame
//     kernel on the GPU: pred= and commit= are their shares of the frame's device time (cycle
//     accounting in the kernel, pbd_b200.h pbd_step_stats), solve= the rest.
//   * by default the server accepts the next client after a disconnect; --once restores the
//     reference's "one client per process lifetime" (Net.cpp:82-93).
//
// There is no CPU fallback: without a CUDA device MSG_INIT fails and the connection is closed.
#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <sys/types.h>
#include <sys/uio.h>
#include <unistd.h>

#include <cerrno>
#include <chrono>
#include <csignal>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pbd_b200.h"

namespace {

constexpr uint32_t kMagic = 0x31444250u;   // 'PBD1', little-endian
enum : uint32_t { MSG_INIT = 1, MSG_STEP = 2, MSG_POSITIONS = 3, MSG_SHUTDOWN = 4 };
constexpr uint64_t kMaxPayload = 1ull << 32;   // the header's size field is 32 bits

#pragma pack(push, 1)
struct Header {
  uint32_t magic, type, size;
};
#pragma pack(pop)
static_assert(sizeof(Header) == 12, "PBD1 header is 12 packed bytes");

bool recv_all(int fd, void* dst, size_t n) {
  char* p = static_cast<char*>(dst);
  while (n) {
    const ssize_t r = ::recv(fd, p, n, 0);
    if (r == 0) return false;
    if (r < 0) { if (errno == EINTR) continue; return false; }
    p += r;
    n -= (size_t)r;
  }
  return true;
}

// header + payload in one gathered send (partial sends are continued)
bool send_message(int fd, uint32_t type, const void* payload, size_t bytes) {
  Header h{kMagic, type, (uint32_t)bytes};
  iovec iov[2] = {{&h, sizeof h}, {const_cast<void*>(payload), bytes}};
  int cnt = bytes ? 2 : 1;
  iovec* cur = iov;
  while (cnt) {
    msghdr m{};
    m.msg_iov = cur;
    m.msg_iovlen = (size_t)cnt;
    ssize_t w = ::sendmsg(fd, &m, MSG_NOSIGNAL);
    if (w < 0) { if (errno == EINTR) continue; return false; }
    while (w > 0 && cnt) {
      if ((size_t)w >= cur->iov_len) { w -= (ssize_t)cur->iov_len; ++cur; --cnt; }
      else { cur->iov_base = static_cast<char*>(cur->iov_base) + w; cur->iov_len -= (size_t)w; w = 0; }
    }
  }
  return true;
}

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

struct Args {
  int port = 7777, device = 0;
  bool once = false;
  pbd_options opts{};
};

void usage(const char* exe) {
  std::printf("Usage:\n  %s --port 7777 [--device N] [--order strict|interleaved|riding] [--fast] [--tagged] [--once]\n", exe);
}

bool parse(int argc, char** argv, Args& a) {
  a.opts.struct_size = sizeof(a.opts);
  a.opts.order_mode = PBD_ORDER_STRICT;   // opts default = the reference's sweep order (edges, then tets)
  for (int i = 1; i < argc; ++i) {
    const std::string s = argv[i];
    if (s == "--help" || s == "-h") { usage(argv[0]); std::exit(0); }
    else if (s == "--port" && i + 1 < argc) a.port = std::atoi(argv[++i]);
    else if (s == "--device" && i + 1 < argc) a.device = std::atoi(argv[++i]);
    else if (s == "--mode" && i + 1 < argc) { if (std::string(argv[++i]) != "gpu") { std::fprintf(stderr, "Unknown mode (this server only has --mode gpu)\n"); return false; } }
    else if (s == "--order" && i + 1 < argc) {
      const std::string o = argv[++i];
      if (o == "strict") a.opts.order_mode = PBD_ORDER_STRICT;
      else if (o == "interleaved") a.opts.order_mode = PBD_ORDER_INTERLEAVED;
      else if (o == "riding") a.opts.order_mode = PBD_ORDER_RIDING;
      else { std::fprintf(stderr, "Unknown order: %s\n", o.c_str()); return false; }
    }
    else if (s == "--fast") a.opts.flags |= PBD_FLAG_FAST_ARITH;
    else if (s == "--tagged") a.opts.flags |= PBD_FLAG_TAGGED_HANDOVER;
    else if (s == "--once") a.once = true;
    else { bool num = !s.empty(); for (char c : s) num &= c >= '0' && c <= '9'; if (num) a.port = std::atoi(s.c_str()); else { std::fprintf(stderr, "Unknown arg: %s\n", s.c_str()); return false; } }
  }
  return true;
}

// one client session; returns false when the server should exit (MSG_SHUTDOWN)
bool serve(int fd, const Args& a) {
  int one = 1;
  ::setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
  pbd_handle* h = nullptr;
  pbd_info info{};
  std::vector<unsigned char> payload;
  std::vector<float> pos;
  bool keepServing = true;
  int frames = 0;
  pbd_step_stats acc{};
  double lastPrint = now_ms();
  uint32_t substeps = 1, iterations = 0;

  for (;;) {
    Header hd{};
    if (!recv_all(fd, &hd, sizeof hd) || hd.magic != kMagic) break;                    // Server.cpp:4-8,25
    if ((uint64_t)hd.size > kMaxPayload) break;
    try { payload.resize(hd.size); } catch (...) { std::fprintf(stderr, "[PBDServer] payload of %u bytes refused (out of memory)\n", hd.size); break; }
    if (hd.size && !recv_all(fd, payload.data(), hd.size)) break;                      // Server.cpp:27-28

    if (hd.type == MSG_INIT) {
      pbd_destroy(h);                                                                   // a second INIT replaces the body (Server.cpp:106-110)
      int st = PBD_OK;
      h = pbd_create_from_init(payload.data(), payload.size(), a.device, &a.opts, &st);
      if (!h) {
        std::fprintf(stderr, "[PBDServer] Init refused (%d): %s\n", st, pbd_last_error());
        break;                                                                          // reference convention: break the loop and shut the session down
      }
      pbd_get_info(h, &info);
      uint32_t pinned = 0;
      std::memcpy(&pinned, payload.data() + 60, 4);
      std::memcpy(&substeps, payload.data() + 12, 4);
      std::memcpy(&iterations, payload.data() + 16, 4);
      if (substeps < 1) substeps = 1;
      pos.assign(3 * (size_t)info.V, 0.0f);
      std::printf("[PBDServer] Init received. V=%u E=%u T=%u pinned=%u  (plan %.0f ms, upload %.0f ms, backend %s)\n", info.V, info.E,
                  info.T, pinned, info.plan_ms, info.upload_ms, pbd_backend_name(h));
      std::fflush(stdout);
    } else if (hd.type == MSG_STEP) {
      if (hd.size < sizeof(float)) break;                                               // Server.cpp:116
      if (!h) continue;                                                                 // STEP before INIT: ignored, no reply (Server.cpp:122)
      float dt;
      std::memcpy(&dt, payload.data(), sizeof dt);
      pbd_step_stats st{};
      if (pbd_step(h, dt, &st) != PBD_OK || pbd_read_positions(h, pos.data(), &st.packMs) != PBD_OK) {
        std::fprintf(stderr, "[PBDServer] step failed: %s\n", pbd_last_error());
        break;
      }
      if (!send_message(fd, MSG_POSITIONS, pos.data(), pos.size() * sizeof(float))) break;   // size == 12 V exactly (PBDRemoteWorld.cs:230-231)
      ++frames;
      acc.predictMs += st.predictMs; acc.solveMs += st.solveMs; acc.commitMs += st.commitMs; acc.packMs += st.packMs; acc.totalMs += st.totalMs;
      const double now = now_ms();
      if (now - lastPrint >= 1000.0) {                                                  // the reference's 1 Hz line (Sim.cpp:400-421), extended
        const double fps = frames * 1000.0 / (now - lastPrint), n = frames;
        const double subPerS = fps * substeps;
        std::printf("[PBDServer] Mode=%s FPS %.1f | V=%u E=%u T=%u | avg(ms): total=%.3f pred=%.3f solve=%.3f commit=%.3f pack=%.3f"
                    " | substeps/s=%.0f tet-constraints/s=%.3g GB/s=%.1f (pred/solve/commit: shares of one fused frame kernel)\n",
                    pbd_backend_name(h), fps, info.V, info.E, info.T, acc.totalMs / n, acc.predictMs / n, acc.solveMs / n, acc.commitMs / n,
                    acc.packMs / n, subPerS, subPerS * (double)info.T * iterations,
                    acc.solveMs > 0 ? (double)info.algorithmic_bytes_per_substep * substeps * n / ((acc.predictMs + acc.solveMs + acc.commitMs) * 1e-3) / 1e9 : 0.0);
        std::fflush(stdout);
        frames = 0;
        acc = pbd_step_stats{};
        lastPrint = now;
      }
    } else if (hd.type == MSG_SHUTDOWN) {
      keepServing = false;                                                              // Server.cpp:138-140
      break;
    } else {
      break;                                                                            // unknown type (Server.cpp:141-143)
    }
  }
  pbd_destroy(h);
  return keepServing;
}

}  // namespace

int main(int argc, char** argv) {
  Args a;
  if (!parse(argc, argv, a)) { usage(argv[0]); return 1; }
  std::signal(SIGPIPE, SIG_IGN);
  const int srv = ::socket(AF_INET, SOCK_STREAM, 0);
  if (srv < 0) { std::perror("socket"); return 1; }
  int one = 1;
  ::setsockopt(srv, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
  sockaddr_in addr{};
  addr.sin_family = AF_INET;
  addr.sin_addr.s_addr = htonl(INADDR_ANY);
  addr.sin_port = htons((uint16_t)a.port);
  if (::bind(srv, reinterpret_cast<sockaddr*>(&addr), sizeof addr) < 0 || ::listen(srv, 1) < 0) { std::perror("bind/listen"); return 1; }
  socklen_t len = sizeof addr;
  ::getsockname(srv, reinterpret_cast<sockaddr*>(&addr), &len);
  std::printf("[PBDServer] Start. mode=gpu devices=%d device=%d port=%d\n", pbd_device_count(), a.device, (int)ntohs(addr.sin_port));
  std::printf("[PBDServer] Listening on port %d...\n", (int)ntohs(addr.sin_port));
  std::fflush(stdout);
  for (;;) {
    const int fd = ::accept(srv, nullptr, nullptr);
    if (fd < 0) { if (errno == EINTR) continue; std::perror("accept"); break; }
    std::printf("[PBDServer] Client connected.\n");
    std::fflush(stdout);
    const bool more = serve(fd, a);
    ::close(fd);
    std::printf("[PBDServer] Client disconnected.\n");
    std::fflush(stdout);
    if (!more || a.once) break;
  }
  ::close(srv);
  std::printf("[PBDServer] Shutdown.\n");
  return 0;
}
