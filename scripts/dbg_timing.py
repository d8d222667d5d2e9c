"""Per-stage clock64 trace of one epilogue warp (TDM_DBG_TIMING build): python scripts/dbg_timing.py M N K tile_n pair
tags: 1 tile start, 2 accumulator ready, 3 chunk in registers, 4 ring slot free, 5 staged, 6 fenced, 7 store issued"""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
from tdm_b200 import ops, _lib
M, N, K, tn, pr = (int(x) for x in sys.argv[1:6])
a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
L = _lib.lib()
buf = (C.c_longlong * 4096)()
for _ in range(3):
    ops.gemm(a, w, out_bf16=o, tile_n=tn, cta_pair=pr)
    L.tdm_dbg_read(buf, 4096)
ev = sorted((buf[i], i % 8, i // 8) for i in range(4096) if buf[i] > 0)
t0 = ev[0][0]
prev = t0
for t, tag, cj in ev[:160]:
    print(cj, tag, t - t0, t - prev)
    prev = t
