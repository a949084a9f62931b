"""Which requests do the reference's own example / tutorial / benchmark scripts make of a binder, and does this engine accept
every one of them?  (Build container only: needs /root/reference; no GPU.)

Each script runs in a subprocess with the unmodified reference importable, plotting modules stubbed out, and a recording
stand-in at ``keops_binder["cpp"]`` (CPU tensors are routed there): for every new (formula, aliases, dtype, flags) it
constructs ``LoadKeOpsB200`` as the GPU request it would be on the box -- parse, autodiff, lowering, kernel generation; with
``--compile`` also NVRTC for sm_100a into the in-tree cache -- and then lets the reference's own CPU binder compute the values
so that the script carries on.  Prints one JSON line per script and a summary.

    python experiments/reference_examples_acceptance.py [--compile] [--only SUBSTR] [--timeout S]
"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

CHILD = r'''
import json, os, runpy, sys, types
from unittest import mock
for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors", "matplotlib.animation", "imageio",
             "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.axes_grid1", "seaborn", "plotly", "plotly.graph_objs",
             "si_prefix"):
    sys.modules[name] = mock.MagicMock()
sys.path.insert(0, os.path.dirname(os.path.abspath(sys.argv[1])))      # sibling helper modules (benchmarks/benchmark_utils.py)
import torch
import pykeops
from pykeops.common import keops_io
from keops_b200 import binder as kb
from keops_b200.runtime import plan
real = keops_io.keops_binder["cpp"]
seen, rejected = set(), []
def shim(tagCPUGPU, tag1D2D, tagHostDevice, use_ranges, device_id, formula, aliases, nargs, dtype, lang, flags):
    key = (formula, tuple(aliases), nargs, dtype, tuple(sorted(flags.items())), bool(use_ranges))
    if key not in seen:
        seen.add(key)
        try:
            b = kb.LoadKeOpsB200(1, 0, 1, use_ranges, -1, formula, list(aliases), nargs, dtype, lang, dict(flags))
            if os.environ.get("KB200_ACCEPT_COMPILE") and not b.ks.is_zero:
                plan.get_cubin(b.ks, bool(b.params.use_fast_math) and dtype == "float32")
            print("ACCEPTED " + formula[:200], flush=True)
        except Exception as e:
            rejected.append([formula[:300], list(aliases), dtype, repr(e)[:300]])
            print("REJECTED " + json.dumps(rejected[-1]), flush=True)
    return real(tagCPUGPU, tag1D2D, tagHostDevice, use_ranges, device_id, formula, aliases, nargs, dtype, lang, flags)
shim.reset = getattr(real, "reset", lambda *a, **k: None)
keops_io.keops_binder["cpp"] = shim
status = "ran"
try:
    sys.argv = [sys.argv[1]]
    runpy.run_path(sys.argv[0], run_name="__main__")
except SystemExit:
    pass
except BaseException as e:
    status = "script stopped: " + repr(e)[:200]
print("ACCEPT" + json.dumps({"status": status, "requests": len(seen), "rejected": rejected}))
'''


def main(argv):
    only = argv[argv.index("--only") + 1] if "--only" in argv else ""
    timeout = int(argv[argv.index("--timeout") + 1]) if "--timeout" in argv else 900
    scripts = []
    for sub in ("examples/pytorch", "examples/numpy", "tutorials/*", "tutorials/*/*", "benchmarks"):
        scripts += sorted(glob.glob(os.path.join(REF, "pykeops", "pykeops", sub, "*.py")))
    scripts = [s for s in scripts if only in s and not os.path.basename(s).startswith("__")]
    cache = os.environ.get("KEOPS_CACHE_FOLDER", "/tmp/keops_cache")
    os.makedirs(cache, exist_ok=True)
    env = dict(os.environ, CXX="/usr/bin/g++", KEOPS_CACHE_FOLDER=cache, PYTHONDONTWRITEBYTECODE="1",
               PYTHONPATH=os.pathsep.join((REF + "/keopscore", REF + "/pykeops", ROOT)))
    if "--compile" in argv:
        env["KB200_ACCEPT_COMPILE"] = "1"
    total = {"scripts": 0, "requests": 0, "rejected": 0, "not_run": 0}
    for s in scripts:
        rel = os.path.relpath(s, REF)
        try:
            r = subprocess.run([sys.executable, "-c", CHILD, s], env=env, cwd="/tmp", capture_output=True, text=True,
                               timeout=timeout)
            lines = [l for l in r.stdout.splitlines() if l.startswith("ACCEPT")]
            rec = json.loads(lines[-1][6:]) if lines else {"status": "no report: " + r.stderr.strip()[-200:], "requests": 0,
                                                           "rejected": []}
        except subprocess.TimeoutExpired as e:
            out = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            rej = [json.loads(l[9:]) for l in out.splitlines() if l.startswith("REJECTED ")]
            rec = {"status": "timeout after %d s (requests up to then)" % timeout,
                   "requests": sum(l.startswith(("ACCEPTED ", "REJECTED ")) for l in out.splitlines()), "rejected": rej}
        rec["script"] = rel
        print(json.dumps(rec), flush=True)
        total["scripts"] += 1
        total["requests"] += rec["requests"]
        total["rejected"] += len(rec["rejected"])
        total["not_run"] += rec["requests"] == 0
    print("SUMMARY " + json.dumps(total))


if __name__ == "__main__":
    main(sys.argv[1:])
