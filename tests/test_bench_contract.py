"""bench.py's fixed points (no GPU): the algorithmic FLOP counts of SURVEY.md 8d, the workloads of BASELINE.json, and the
rule that the ONE JSON line goes to the process's original stdout while native banners written to descriptor 1 do not."""
import importlib.util
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_needed_flops_match_the_survey():
    b = _bench()
    assert abs(b.needed_flops_per_image("vits8", 4097) / 4.449e11 - 1) < 1e-3
    assert abs(b.needed_flops_per_image("vitb8", 4097) / 1.211e12 - 1) < 1e-3
    assert {"cfg1", "cfg2", "cfg3"} <= set(b.WORKLOADS)
    size, arch, fos, n_cls, _, _ = b.WORKLOADS["cfg2"]
    assert (size, arch, fos, n_cls) == (256, "vits8", 64, 8)                        # BASELINE.json configs[1]
    size, arch, fos, n_cls, _, _ = b.WORKLOADS["cfg3"]
    assert (size, arch, fos, n_cls) == (512, "vitb8", 64, 16)                       # configs[2]


def test_json_line_is_alone_on_stdout():
    code = (
        "import os, sys, importlib.util\n"
        f"spec = importlib.util.spec_from_file_location('b', r'{ROOT / 'bench.py'}')\n"
        "b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)\n"
        "sys.stdout.flush(); b._JSON_FD = os.dup(1); os.dup2(2, 1)\n"          # what main() does first
        "os.write(1, b'NCCL version banner\\n')\n"                             # a native library writing to fd 1
        "print('python-level noise')\n"
        "b._emit({'metric': 'x', 'value': 1.0})\n"
    )
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ))
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1 and json.loads(lines[0]) == {"metric": "x", "value": 1.0}
    assert "NCCL version banner" in r.stderr and "python-level noise" in r.stderr
