"""The reference's OWN main.py drives this repository unchanged: a byte-identical
copy placed outside the reference tree, and the file in place under PYTHONSAFEPATH=1.  Two ranks, gloo CPU
plumbing configuration (BASELINE config 1: Vanilla, fp32), synthetic partitions; the run must finish and
write the reference's exp/ files (trainer.py:203-238).  Needs /root/reference (the build container)."""
import os
import shutil
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_MAIN = "/root/reference/main.py"


def _port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(not os.path.exists(REF_MAIN), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("form", ["copy", "safepath"])
def test_reference_main_runs_unmodified(form, tmp_path):
    env = dict(os.environ)
    env.update({"PYTHONPATH": ROOT, "ADAQP_DEVICE": "cpu", "ADAQP_SYNTHETIC": "1", "ADAQP_SYNTH_SCALE": "0.003",
                "ADAQP_NUM_EPOCHES": "3", "ADAQP_SEED": "5", "OMP_NUM_THREADS": "1"})
    if form == "copy":
        script = str(tmp_path / "main_ref.py")
        shutil.copyfile(REF_MAIN, script)
        assert open(script, "rb").read() == open(REF_MAIN, "rb").read()
    else:
        script = REF_MAIN
        env["PYTHONSAFEPATH"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_port()), script, "--dataset", "ogbn-products", "--num_parts", "2", "--model_name", "gcn",
           "--mode", "Vanilla", "--assign_scheme", "uniform", "--logger_level", "WARNING"]
    r = subprocess.run(cmd, cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    base = tmp_path / "exp" / "ogbn-products" / "2part" / "gcn"
    assert (base / "time" / "Vanilla.csv").exists() and (base / "metrics" / "Vanilla.txt").exists()
    rows = (base / "time" / "Vanilla.csv").read_text().strip().splitlines()
    assert rows[0].split(",")[:3] == ["Worker", "Overhead", "Total"] and len(rows) == 3
