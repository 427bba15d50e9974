"""Analytic bounds for the hot kernels at the BASELINE configurations (no GPU needed).

For every configuration: algorithmic work, the time the measured peaks of this pool's
B200s allow (MEASURED_PEAKS.json, else the fallback constants of B200_PROFILING.md), and — for the
tcgen05 kernels — the per-tile bounds that the tile geometry itself imposes (tensor pipe, shared-memory
operand reads + TMA writes at 128 B/clk/SM, L2 -> SM operand stream at ~6300 B/clk chip-wide).
Writes a markdown table; measured numbers are quoted from profiles/ for comparison."""
import json
import os

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
try:
    PK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    TF_BURST, TF_SUST, HBM, SRC = PK["bf16_tflops"], PK["bf16_tflops_sustained"], PK["hbm_gbs"], "MEASURED_PEAKS.json"
except Exception:
    TF_BURST, TF_SUST, HBM, SRC = 1689.8, 1402.5, 6568.4, "fallback constants"
SMS, SMEM_BPC, L2_BPC = 148, 128.0, 6300.0


def tile_bounds(bm, bn, k_elems, a_resident):
    """Cycles per (bm x bn) output tile of an SS-mode bf16 MMA pipeline with M = 128 instructions."""
    halves = bm // 128
    n_mma = halves * (k_elems // 16)
    tensor = n_mma * 128 * bn / 256.0                       # max(M,128) * N / 256 cycles per instruction
    smem_reads = n_mma * (128 * 16 * 2 + bn * 16 * 2)       # A slice + B slice per instruction
    tma_writes = bn * k_elems * 2 + (0 if a_resident else bm * k_elems * 2)
    smem = (smem_reads + tma_writes) / SMEM_BPC
    l2 = tma_writes / (L2_BPC / SMS)
    return tensor, smem, l2


rows = []
# C2: DSA 10k x 60k x 128 (K = 144 packed), resident-query kernel 256 x 192
pairs2 = 5.4e8
for name, pairs in (("C2 DSA stage 2 (other classes)", 5.4e8), ("C2 DSA stage 1 (same class)", 6.0e7)):
    flops = 2 * 128 * pairs
    tensor, smem, l2 = tile_bounds(256, 192, 144, True)
    tiles = pairs / (256 * 192) / SMS
    rows.append((name, f"{flops:.3g} flop", f"{flops / TF_BURST / 1e12 * 1e6:.1f} us @ {TF_BURST:.0f} TF/s",
                 f"tile 256x192, K=144: tensor {tensor:.0f} / smem {smem:.0f} / L2 {l2:.0f} cycles -> "
                 f"{max(tensor, smem, l2) * tiles / 1.66e3:.1f} us at 1.66 GHz for {tiles:.0f} tiles per SM",
                 "0.117-0.126 ms (stage 2), 0.034 ms (stage 1)" if "stage 2" in name else "see above"))
# C3: LSA 10k x 60k x 256, three segments (K = 784), streaming kernel 128 x 256
pairs3 = 6.0e8
flops3 = 2 * 256 * pairs3
tensor, smem, l2 = tile_bounds(128, 256, 784, False)
tiles = pairs3 / (128 * 256) / SMS
rows.append(("C3 LSA (3-segment split-bf16)", f"{flops3:.3g} flop algorithmic, {2 * 784 * pairs3:.3g} executed",
             f"{flops3 / TF_BURST / 1e12 * 1e6:.0f} us algorithmic, {2 * 784 * pairs3 / TF_BURST / 1e12 * 1e6:.0f} us executed",
             f"tile 128x256, K=784: tensor {tensor:.0f} / smem {smem:.0f} / L2 {l2:.0f} cycles -> "
             f"{max(tensor, smem, l2) * tiles / 1.66e3:.0f} us at 1.66 GHz (L2-bound)", "0.84 ms whole device step"))
tensor, smem, l2 = tile_bounds(256, 256, 784, False)
tiles = pairs3 / (256 * 256) / SMS
rows.append(("C3 LSA with 256x256 tiles (next round)", "same", "same",
             f"tensor {tensor:.0f} / smem {smem:.0f} / L2 {l2:.0f} cycles -> {max(tensor, smem, l2) * tiles / 1.66e3:.0f} us",
             "not built"))
# C4 KMNC, C1 DeepGini
b4 = 10000 * 4096 * (4 + 2) + 2 * 4096 * 4 + 10000 * 4
rows.append(("C4 KMNC 10k x 4096 x 1000", f"{b4 / 1e6:.0f} MB", f"{b4 / HBM / 1e9 * 1e6:.1f} us @ {HBM:.0f} GB/s", "-",
             "64.5 us per step (3.8 TB/s)"))
b1 = 10000 * 10 * 4 + 10000 * 8
rows.append(("C1 DeepGini 10k x 10", f"{b1 / 1e3:.0f} KB", f"{b1 / HBM / 1e9 * 1e6:.2f} us", "launch latency only", "10 us per call"))
# C5 per GPU (1/8 of the train set), streaming kernel at K = 2064
pairs5 = 1.0e5 * 1.28e6 / 8
flops5 = 2 * 2048 * pairs5
tensor, smem, l2 = tile_bounds(128, 256, 2064, False)
tiles = pairs5 / (128 * 256) / SMS
rows.append(("C5 DSA per GPU (100k x 160k x 2048)", f"{flops5:.3g} flop", f"{flops5 / TF_SUST / 1e12 * 1e3:.1f} ms @ {TF_SUST:.0f} TF/s sustained",
             f"tile 128x256, K=2064: tensor {tensor:.0f} / smem {smem:.0f} / L2 {l2:.0f} cycles -> "
             f"{max(tensor, smem, l2) * tiles / 1.66e6:.1f} ms", "74 ms per pass (first half of the round)"))

out = ["# Analytic bounds per configuration (tools/roofline_model.py)", "",
       f"Peaks: {SRC} — bf16 {TF_BURST:.0f} TF/s burst / {TF_SUST:.0f} sustained, HBM {HBM:.0f} GB/s; per-tile bounds assume "
       "128 B/clk/SM of shared-memory bandwidth (MMA operand reads + TMA writes) and ~6300 B/clk of L2 -> SM bandwidth "
       "chip-wide (B300_MICROARCH), at the 1.66 GHz the SMs hold under the tcgen05 kernels (ncu).", "",
       "| configuration | algorithmic work | time at the measured peak | geometry bounds | measured |", "|---|---|---|---|---|"]
out += [f"| {' | '.join(r)} |" for r in rows]
path = os.path.join(ROOT, "profiles", "r01b_roofline_model.md")
open(path, "w").write("\n".join(out) + "\n")
print("\n".join(out))
