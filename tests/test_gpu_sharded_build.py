"""GPU (-m gpu): lnr_index_build_sharded -- the hash-range sharded DIndex / HIndex build with the NCCL exchange INSIDE the C ABI.
One process per rank (as many ranks as GPUs on the box, at most 4; a single GPU runs the 1-rank form of the same code
path). Every rank must end up with the oracle's dir / hs and map the case's reads to the oracle's cords."""
import os
import subprocess
import sys
import tempfile

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("case_name,index_type", [("repeat_ont", 1), ("repeat_t16", 1), ("repeat_ont", 2)])
def test_sharded_build_inside_the_c_abi(case_name, index_type):
    import torch
    n_gpu = torch.cuda.device_count()
    world = 4 if n_gpu >= 4 else (2 if n_gpu >= 2 else 1)
    with tempfile.TemporaryDirectory() as td:
        idfile = os.path.join(td, "nccl_id")
        procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "sharded_worker.py"), str(r), str(world), idfile, case_name, str(index_type)],
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
        outs = []
        for p in procs:
            try:
                out, _ = p.communicate(timeout=600)
            except subprocess.TimeoutExpired:
                for q in procs:
                    q.kill()
                raise
            outs.append(out)
        for r, (p, out) in enumerate(zip(procs, outs)):
            assert p.returncode == 0 and ("rank %d ok" % r) in out, out[-2000:]
