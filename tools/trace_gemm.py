"""Timeline of CTA 0 for one training-path GEMM (debug aid): clock64 stamps of the MMA warp (stream 0) and two worker
warps (streams 1, 2).  Tags: 1/2 first/later chunk ready, 3 item fully issued; producers: 50/51 item begin/end, per chunk 52
loads issued, 53 ring stage free, 54 converted + stored, 55 arrived; drain: 10/11 wait for / got the accumulators, per
32-column block 13 TMEM load done, 14 staged, 12 stored; 61 item done.  ANERF_TC_DEBUG (bit 0: no C stores in the forward
form, bit 1: no A loads, bit 2: 16-byte B copies) switches traffic off for timing experiments (profiles/r2_tc_gemm_analysis.md).

    python tools/trace_gemm.py [rows] [N] [K] [form: fwd|dgrad|wgrad]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from anerf_b200 import _lib
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 196608
N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
K = int(sys.argv[3]) if len(sys.argv) > 3 else 256
form = sys.argv[4] if len(sys.argv) > 4 else "fwd"
dev = torch.device("cuda")
torch.manual_seed(0)
X = torch.randn(rows, K, device=dev)
W = torch.randn(N, K, device=dev) / K ** 0.5
b = torch.randn(N, device=dev)
C = torch.empty(rows, N, device=dev)
G = torch.randn(rows, N, device=dev)
dW = torch.zeros(N, K, device=dev)


def run():
    if form == "fwd":
        _lib.selftest_tc_gemm(X, (K, 1), rows, K, W, (K, 1), N, C, (N, 1), bias=b, relu=True)
    elif form == "dgrad":      # dX[rows, K] = G[rows, N] W[N, K], masked by X
        _lib.selftest_tc_gemm(G, (N, 1), rows, N, W, (1, K), K, X, (K, 1), mask=C, mask_ms=N if N == K else 0)
    else:                      # dW[n, k'] += sum_rows G[row, n] X[row, k']
        _lib.selftest_tc_gemm(X, (1, K), K, rows, G, (1, N), N, dW, (1, K), mode=2, slice_chunks=32)


for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print("call ms (pack + gemm + sync)", e0.elapsed_time(e1))
buf = torch.zeros(3 * 1024, dtype=torch.int64, device=dev)
_lib.load().anerf_debug_set_trace(buf.data_ptr())
run()
torch.cuda.synchronize()
_lib.load().anerf_debug_set_trace(None)
tb = buf.cpu().numpy().reshape(3, 1024)
t0 = None
for s in range(3):
    n = int(tb[s, 1023]); ev = [(int(x >> 48), int(x & 0xFFFFFFFFFFFF)) for x in tb[s, :n]]
    if not ev:
        continue
    if t0 is None: t0 = ev[0][1]
    print(f"--- stream {s}: {n} events")
    prev = None; line = []
    for tag, c in ev[:150]:
        line.append(f"{tag}@{c - t0}" + (f"(+{c - prev})" if prev is not None else ""))
        prev = c
    print(" ".join(line))
