"""Device-side timing of a launch sequence without host overhead: the calls are captured ONCE into a CUDA graph
(`n` launches over rotating buffer sets) and the graph is replayed -- a Python / ctypes / tensor-map-encode call costs
40-80 us on the host, more than most single launches take on the GPU."""
import torch


def graph_time_us(fn, n=10, replays=5):
    """fn(i) issues launch i (i selects the buffer set).  Returns the average microseconds per launch."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for i in range(3):
            fn(i)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        for i in range(n):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (n * replays)
