"""Every opt-in kernel variant selected by an environment variable must give the same results as the
default path: the TMA bulk-copy ring of the dense batch kernel, the unstaged sequential CSR kernel,
the exhaustive float64 line search and the search without the float32 refinement stage."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _run(env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    r = subprocess.run([sys.executable, os.path.join(HERE, "_variant_probe.py")], env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("DIGEST ")][-1]
    return json.loads(line[len("DIGEST "):])


@pytest.mark.gpu
def test_kernel_variants_agree():
    base = _run({})
    for env, keys in (({"XCOLUMNS_B200_DENSE_PATH": "tma"}, ["bca_batched"]),
                      ({"XCOLUMNS_B200_DENSE_R": "2"}, ["bca_batched", "fw"]),
                      ({"XCOLUMNS_B200_EXACT_CSR": "generic"}, ["bca_exact_csr"]),
                      ({"XCOLUMNS_B200_FW_SEARCH": "full"}, ["fw"]),
                      ({"XCOLUMNS_B200_FW_REFINE": "0"}, ["fw"])):
        got = _run(env)
        for key in keys:
            assert got[key] == base[key], (env, key, got[key], base[key])
