#!/usr/bin/env python3
"""Copy the ncu summaries a tools/gpu_r2_final*.sh run left in gpurun_out/ into profiles/ (tracked) and rebuild
profiles/traffic.json (what bench.py reports as roofline.traffic / issue_slot_frac / smem_pipe_frac) from them."""
import json, os, re, shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
CAPS = {
    "tile_fast": ("r2_tile_frame_kernel_fast_ncu_full.txt", "headline:b200-tile-tagged-fast",
                  "ncu --set full --clock-control none --import-source on -k regex:tile_frame -s 2 -c 1  (python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sustained --no-extra --arith fast)\n"
                  "kernel: pbd::tile_frame_kernel<1, TAGGED, FAST, no inert multipliers, RESIDENT record blocks>, grid 294 x 256 threads (two tiles per SM), one launch = one frame = 20 substeps x 6 iterations of the 1M-tet Kuhn grid (interleaved order, mixed colour steps, tagged 128-bit hand-over, fast arithmetic, placement search) -- the `value` configuration of bench.py; final round-2 build\n"),
    "tile_exact": ("r2_tile_frame_kernel_exact_ncu_full.txt", "headline:b200-tile-tagged",
                   "ncu --set full --clock-control none --import-source on -k regex:tile_frame -s 2 -c 1  (python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sustained --no-extra --arith exact)\n"
                   "kernel: pbd::tile_frame_kernel<1, TAGGED, exact>, grid 294 x 256 threads, same workload, bit-exact arithmetic -- the `alt.exact` configuration of bench.py; final round-2 build\n"),
    "batch_fast": ("r2_batch_frame_kernel_fast_ncu_full.txt", "batch4096:b200-batch-fast",
                   "ncu --set full --clock-control none --import-source on -k regex:batch_frame -s 2 -c 1  (python bench.py --workload batch4096 --steps 1 --warmup 3 --no-cpu-baseline --arith fast)\n"
                   "kernel: batch_frame_kernel<1, FAST>, 148 x 512 threads, one launch = one frame = 10 substeps x 6 iterations of 4096 x 6k-tet bodies; final round-2 build\n"),
    "batch_exact": ("r2_batch_frame_kernel_exact_ncu_full.txt", "batch4096:b200-batch",
                    "ncu --set full --clock-control none --import-source on -k regex:batch_frame -s 2 -c 1  (python bench.py --workload batch4096 --steps 1 --warmup 3 --no-cpu-baseline --arith exact)\n"
                    "kernel: batch_frame_kernel<1, exact>, 148 x 512 threads, same workload, bit-exact arithmetic; final round-2 build\n"),
}


def parse(path):
    out = {}
    for line in open(path):
        m = re.match(r"(\S+) \[(.*?)\] = (\S+)", line)
        if m:
            out[m.group(1)] = float(m.group(3)) * UNIT.get(m.group(2), 1.0)
    return out


def main():
    tj = os.path.join(P, "traffic.json")
    traffic = json.load(open(tj))
    for name, (dst, key, head) in CAPS.items():
        src = os.path.join(G, f"r2_{name}_summary.txt")
        if not os.path.exists(src):
            print("missing", src)
            continue
        body = open(src).read()
        open(os.path.join(P, dst), "w").write(head + body)
        m = parse(src)
        traffic[key] = {
            "dram_bytes": m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"],
            "warp_instructions": m["smsp__inst_executed.sum"],
            "smem_wavefronts": m["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"],
            "smem_bank_conflict_wavefronts": m["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"],
            "kernel_ms_under_ncu": m["gpu__time_duration.sum"],
            "source": "profiles/" + dst,
        }
        print(key, traffic[key])
    json.dump(traffic, open(tj, "w"), indent=1)
    ll = os.path.join(G, "r2_launches.csv")
    if os.path.exists(ll):
        shutil.copy(ll, os.path.join(P, "r2_launches.csv"))


if __name__ == "__main__":
    main()
