"""Graph-timed launch helper shared by the tc32 probes."""
import torch


def timeit(fn, n=20):
    """fn(stream) launched n times inside ONE CUDA graph, replayed five times; us per launch."""
    for _ in range(3):
        fn(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(n):
            fn(st)
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (5 * n) * 1e3
