#!/usr/bin/env python
"""Acting-forward throughput of the (unchanged) DQN net on one GPU: fp32 / bf16 autocast / bf16 weights, NCHW / channels_last."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tron_b200  # noqa: E402
from tron_b200 import dropin  # noqa: E402
dropin.install()
from Net.DQNNet import Net  # noqa: E402


def timed(fn, iters=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    rows, planes = int(sys.argv[1]) if len(sys.argv) > 1 else 262144, int(sys.argv[2]) if len(sys.argv) > 2 else 3
    net = Net(planes).cuda().eval()
    x = torch.randint(0, 2, (rows, planes, 12, 12), device="cuda").to(torch.bfloat16)
    out = {}
    with torch.no_grad():
        out["fp32_nchw"] = timed(lambda: net(x.float()))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out["autocast_bf16_nchw"] = timed(lambda: net(x))
        nb = Net(planes).cuda().eval().to(torch.bfloat16)
        out["bf16_weights_nchw"] = timed(lambda: nb(x))
        nbc = Net(planes).cuda().eval().to(torch.bfloat16).to(memory_format=torch.channels_last)
        xc = x.contiguous(memory_format=torch.channels_last)
        out["bf16_weights_channels_last_incl_input_conversion"] = timed(lambda: nbc(x.contiguous(memory_format=torch.channels_last)))
        out["bf16_weights_channels_last"] = timed(lambda: nbc(xc))
        torch.backends.cudnn.benchmark = True
        out["bf16_weights_channels_last_cudnn_benchmark"] = timed(lambda: nbc(xc))
        out["bf16_weights_nchw_cudnn_benchmark"] = timed(lambda: nb(x))
        for chunk in (32768, 65536):
            out["bf16_weights_channels_last_chunks_%d" % chunk] = timed(lambda: [nbc(xc[i:i + chunk]) for i in range(0, rows, chunk)])
    flops = rows * 36.2e6
    print(json.dumps({"rows": rows, "planes": planes, "ms": out, "TFLOPs": {k: flops / (v * 1e-3) / 1e12 for k, v in out.items()}}))


if __name__ == "__main__":
    main()
