"""Generate tests/golden/relpos_table.npz by RUNNING the reference's RelativePositionalEmbeddings
(/root/reference/unused_lib/relativepositionalembeddings.py, staged as oracle/_ref/lib/relativepositionalembeddings.py by
the R8 shim).  Run once in the build container:  python tests/golden/make_golden_relpos.py"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import stage_ref  # noqa: E402

assert stage_ref.stage(), "reference not available"
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
from lib.relativepositionalembeddings import RelativePositionalEmbeddings  # noqa: E402

out = {}
for shape in [(12, 16), (80, 256), (3, 5, 8)]:
    t = RelativePositionalEmbeddings(shape, True)()
    assert t.requires_grad and tuple(t.shape) == shape
    out["t_" + "x".join(map(str, shape))] = t.detach().numpy().copy()
big = RelativePositionalEmbeddings((1280, 512), False)()          # hp.batch_size * max_len = 32 * 40 rows, conformer_size 512
assert not big.requires_grad
out["sha_1280x512"] = np.frombuffer(hashlib.sha256(big.contiguous().numpy().tobytes()).digest(), dtype=np.uint8)
np.savez_compressed(os.path.join(HERE, "relpos_table.npz"), **out)
print({k: v.shape for k, v in out.items()})
