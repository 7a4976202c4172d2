"""Stand-alone check/timing of the quantised GEMM through the C ABI (debug + profiling helper).
  python tools/gemm_check.py [m] [n] [k] [type: 12=Q4_K 14=Q6_K] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import candle_vllm_b200 as pkg
from candle_vllm_b200 import synthetic

if os.environ.get("B200_TRACE"):
    import torch as _t
    _trace = _t.zeros(34 * 8, dtype=_t.int64, device="cuda")
    os.environ["B200_GEMM_TRACE"] = str(_trace.data_ptr())
m = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
k = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
t = int(sys.argv[4]) if len(sys.argv) > 4 else 12
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 0
g = torch.Generator(device="cuda"); g.manual_seed(0)
w = synthetic.random_qtensor(g, t, n, k, "cuda")
x = torch.randn((m, k), device="cuda", generator=g)
mm = pkg.QMatMul(w)
y = mm.forward(x)
torch.cuda.synchronize()
wd = w.dequantize()
ref = (x.half().float().double() @ wd.double().T).float()      # fp16-rounded activations, exact weights
ref32 = (x.double() @ wd.double().T).float()
rel = ((y - ref).norm() / ref.norm()).item()
rel32 = ((y - ref32).norm() / ref32.norm()).item()
def show_trace(tag):
    if not os.environ.get("B200_TRACE"): return
    tr = _trace.cpu().numpy().reshape(34, 8)
    t0 = tr[tr > 0].min()
    print(f"[{tag}] unit: prod_empty_ok | mma_full_ok mma_aready_ok mma_committed | deq_full_ok deq_afree_ok deq_done   (cycles since first stamp)")
    for i in range(16):
        if (tr[i] > 0).any(): print(i, [int(v - t0) if v > 0 else -1 for v in tr[i, :7]])
    print("kernel: entry, setup_done, (seg dequant_done, acc_complete) x3:", [int(v - t0) if v > 0 else -1 for v in tr[32]])
    print("kernel: last_epilogue_done, all_roles_done:", [int(v - t0) if v > 0 else -1 for v in tr[33, :2]])
    _trace.zero_()
show_trace("cold first launch")
print(f"m={m} n={n} k={k} type={t}: rel-fro vs fp16-act ref {rel:.3e}, vs f32 ref {rel32:.3e}, max|y|={y.abs().max().item():.3f}")
if reps:
    import ctypes as C
    from candle_vllm_b200.backend import DType
    L = pkg.lib()
    Lw = 12
    ws = [synthetic.random_qtensor(g, t, n, k, "cuda") for _ in range(Lw)]
    xk4 = torch.empty((m, k), dtype=torch.float16, device="cuda")
    xh = x.half().contiguous()
    st = torch.cuda.Stream()
    yb = torch.zeros((m, n), dtype=torch.float32, device="cuda")
    with torch.cuda.stream(st):
        L.cast(C.c_void_p(xh.data_ptr()), C.c_void_p(xk4.data_ptr()), C.c_int64(xh.numel()), C.c_int32(DType.F16), C.c_int32(DType.F16_K4), C.c_int64(st.cuda_stream))
        slabs = int(os.environ.get("B200_SLABS", "0"))
        L.qmatmul_slab_count.restype = C.c_int32
        ns = int(L.qmatmul_slab_count(C.c_int32(m), C.c_int32(n), C.c_int32(k), C.c_int32(t)))
        ysl = torch.empty((max(ns, 1), m, n), dtype=torch.float32, device="cuda")
        def call(i):
            if slabs:
                L.qmatmul_f16act_slabs(C.c_void_p(xk4.data_ptr()), C.c_void_p(ws[i % Lw].data.data_ptr()), C.c_void_p(ysl.data_ptr()), C.c_int32(ysl.shape[0]),
                                       C.c_int32(m), C.c_int32(n), C.c_int32(k), C.c_int32(t), C.c_int64(st.cuda_stream))
                return
            L.qmatmul_f16act(C.c_void_p(xk4.data_ptr()), C.c_void_p(ws[i % Lw].data.data_ptr()), C.c_void_p(yb.data_ptr()), C.c_int32(m), C.c_int32(n),
                             C.c_int32(k), C.c_int32(t), C.c_int32(1), C.c_int64(st.cuda_stream))
        for i in range(3): call(i)
        st.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for i in range(reps): call(i)
        gr.replay(); st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); gr.replay(); e1.record(st); st.synchronize()
    ms = e0.elapsed_time(e1) / reps
    byts = w.data.numel()
    print(f"  graph of {reps} GEMM launches: {ms*1e3:.2f} us/launch, {byts/ms/1e6:.1f} GB/s weight stream (debug={os.environ.get('B200_GEMM_DEBUG','0')}, slabs={ns if slabs else 0})")
    show_trace("last launch of the warm graph")
