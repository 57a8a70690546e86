"""A/B of the staged copies of `algo.update(host batch)`: copy-engine nodes (cudaMemcpyAsync, D3B_ZEROCOPY_MAX=0) vs
kernel nodes over the pinned buffers (csrc/util.cu: copy_mapped).  End-to-end wall clock per update, alternating the
two settings in one process (graphs re-captured on every switch), c2 and c1, both precisions.
Run on a B200 box: python profiles/r2/zerocopy_probe.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402


def main():
    torch.cuda.set_device(0)
    out = {}
    for name in ("c2", "c1"):
        for prec in ("bf16", "fp32"):
            r = bench.Runner(bench.WORKLOADS[name], 1, 0, 0, prec, False)
            default = type(r.impl).zero_copy_max
            res = {"memcpy": [], "kernel": []}
            for rep in range(3):
                for label, zmax in (("memcpy", 0), ("kernel", default)):
                    r.impl.zero_copy_max = zmax
                    r.impl._graphs_invalidate()
                    e = r.e2e(300)[0]
                    res[label].append(round(e["ms_per_step"] * 1e3, 2))
            out[f"{name}_{prec}"] = {"us_per_update": res,
                                     "best": {k: min(v) for k, v in res.items()}}
            print(name, prec, out[f"{name}_{prec}"], flush=True)
            del r
            torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
