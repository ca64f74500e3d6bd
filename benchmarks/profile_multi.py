"""One training step of the fused multi-aggregation at a footprint ncu can save / restore quickly
(`ncu --set full` backs the device memory up around every replay pass: at the 100 M-edge bench shape one kernel takes minutes).

    ncu --set full --clock-control none --import-source on -k regex:multi_aggr -o gpurun_out/r2_multi \
        python benchmarks/profile_multi.py --nodes 2000000 --edges 20000000
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_graph  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=2_000_000)
    ap.add_argument("--edges", type=int, default=20_000_000)
    ap.add_argument("--feat", type=int, default=256)
    ap.add_argument("--steps", type=int, default=1)
    args = ap.parse_args()
    from pytorch_geometric_b200 import functional as Fn
    from pytorch_geometric_b200.graph import CSRGraph
    dev = torch.device("cuda", 0)
    ei = synth_graph(args.nodes, args.edges, 2, dev)
    g = CSRGraph(ei[0], ei[1], args.nodes, args.nodes)
    g.build_transpose()
    _ = g.t2csr
    del ei
    aggrs = ["mean", "min", "max", "std"]
    x = torch.randn(args.nodes, args.feat, device=dev, requires_grad=True)
    gouts = [torch.randn(args.nodes, args.feat, device=dev) for _ in aggrs]
    for _ in range(args.steps):
        x.grad = None
        torch.autograd.backward(Fn.multi_aggregate(g, x, aggrs), gouts)
    torch.cuda.synchronize()
    print("done", float(x.grad.abs().sum()))


if __name__ == "__main__":
    main()
