"""CUDA-graph capture of a launch-bound call sequence (the Blackwell playbook: "capture launch-bound inner loops in CUDA
graphs").  The hot-path calls of this package enqueue kernels only -- no host synchronisation, no data-dependent host control
flow -- so a whole forward (IPDAE round trip, PPPF_AE forward) can be captured once and replayed."""
import torch


def capture(fn, *example_inputs, warmup=2):
    """Capture `fn(*inputs)` for inputs of the examples' shapes / dtypes.  Returns `run(*inputs)`: copies the inputs into the
    graph's static buffers, replays the graph and returns its static outputs (overwritten by the next call)."""
    static = [t.clone() for t in example_inputs]
    dev = static[0].device
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side), torch.no_grad():      # warm-up outside the capture: weight packing, attribute set-up
        for _ in range(warmup):
            fn(*static)
    torch.cuda.current_stream(dev).wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph), torch.no_grad():
        outs = fn(*static)

    def run(*inputs):
        for s, t in zip(static, inputs):
            s.copy_(t, non_blocking=True)
        graph.replay()
        return outs

    return run
