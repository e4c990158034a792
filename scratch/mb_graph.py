"""How fast does a CUDA graph issue tiny kernels: one dependent chain vs parallel chains?"""
import torch, time
dev = 'cuda'
xs = [torch.zeros(256, device=dev) for _ in range(8)]

def build(n_kernels, n_lanes):
    g = torch.cuda.CUDAGraph()
    lanes = [torch.cuda.Stream() for _ in range(n_lanes)]
    with torch.cuda.graph(g):
        main = torch.cuda.current_stream()
        for s in lanes: s.wait_stream(main)
        for i in range(n_kernels):
            with torch.cuda.stream(lanes[i % n_lanes]):
                xs[i % n_lanes].add_(1.0)
        for s in lanes: main.wait_stream(s)
    return g

for n_lanes in (1, 2, 4, 6):
    g = build(168, n_lanes)
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): g.replay()
    b.record(); torch.cuda.synchronize()
    print('168 tiny kernels over %d lanes: %.1f us per replay, %.2f us per kernel' % (n_lanes, a.elapsed_time(b) * 20, a.elapsed_time(b) * 20 / 168))
