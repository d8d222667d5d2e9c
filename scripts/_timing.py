"""Shared timing helper for the micro-benchmarks: CUDA events on the launching stream, >= 3 warm-up calls, the best of
`reps` timed loops of >= `min_ms` each, and the SM clock NVML reports right after the loop (a throttled or idle-clocked
GPU shows up in the record instead of silently skewing a comparison)."""
import torch

try:
    import pynvml
    pynvml.nvmlInit()
    _H = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device() if torch.cuda.is_available() else 0)
except Exception:  # pragma: no cover
    pynvml = None
    _H = None


def sm_mhz():
    if _H is None:
        return None
    try:
        return pynvml.nvmlDeviceGetClockInfo(_H, pynvml.NVML_CLOCK_SM)
    except Exception:
        return None


def timeit(fn, iters=20, warm=3, reps=3, min_ms=20.0):
    """-> (best microseconds per call, SM MHz seen after the best loop)"""
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    best, clk = None, None
    for _ in range(reps):
        n = iters
        while True:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n):
                fn(i)
            e1.record()
            mhz = sm_mhz()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if ms >= min_ms or n >= 4096:
                break
            n *= 2
        us = ms / n * 1e3
        if best is None or us < best:
            best, clk = us, mhz
    return best, clk


def timeit_graph(fn, n=50, reps=5):
    """Device time per call of a SHORT kernel: n calls captured into a CUDA graph and replayed, so the ~12-35 us Python /
    ctypes cost per call (which bounds a plain timing loop of any kernel shorter than that) is out of the measurement.
    -> (best microseconds per call, SM MHz after the best replay)"""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for i in range(3):
            fn(i)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    best, clk = None, None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        mhz = sm_mhz()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / n
        if best is None or us < best:
            best, clk = us, mhz
    return best, clk
