"""Records the order of calls the UNMODIFIED reference `train_single_env` (d3rlpy/online/iterators.py:99-287) makes on
recording fakes (tests/online_loop_fakes.py), for four configurations.  Run in the build container:
    python tests/golden/make_online_loop_trace.py"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
os.environ.setdefault("WANDB_MODE", "disabled")
from oracle import ref_import  # noqa: E402

ref_import.load()
from d3rlpy.online.iterators import train_single_env  # noqa: E402

import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location("online_loop_fakes", os.path.join(HERE, "..", "online_loop_fakes.py"))
_fakes = importlib.util.module_from_spec(_spec)   # by path: the reference ships its own `tests` package
_spec.loader.exec_module(_fakes)
CONFIGS, run = _fakes.CONFIGS, _fakes.run

out = {}
with tempfile.TemporaryDirectory() as d:
    for name, cfg in CONFIGS.items():
        out[name] = run(train_single_env, cfg, logdir=d, show_progress=False, save_metrics=False, verbose=False,
                        with_timestamp=False, experiment_name=name)
json.dump(out, open(os.path.join(HERE, "online_loop_trace.json"), "w"))
print({k: len(v["trace"]) for k, v in out.items()})
