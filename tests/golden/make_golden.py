"""Regenerates tests/golden/golden.npz + golden_meta.json from the REFERENCE's own code
(oracle/_ref/libfmref.so = /root/reference/src/rtl_fm_player.c compiled unmodified by
oracle/Makefile).  Run in the build container only:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.oracle_py import RefOracle  # noqa: E402
from vectors import B, CASES, CONFIGS, LONG_CASES, long_capture_bytes, make_input, sha  # noqa: E402

arrays, meta = {}, {"cases": {}, "long": {}, "tables": {}, "layout": {}}
for cid, cfg, kind, stream, blocks in CASES:
    iq = make_input(cfg, kind, stream, blocks)
    ref = RefOracle(**CONFIGS[cfg])
    pcm, counts = [], []
    for b in range(blocks):
        p = ref.block(iq[b * B:(b + 1) * B])
        pcm.append(p)
        counts.append(len(p))
    arrays[cid] = np.concatenate(pcm)
    meta["cases"][cid] = {"config": cfg, "kind": kind, "stream": stream, "blocks": blocks, "input_sha256": sha(iq),
                          "counts": counts, "pcm_sha256": sha(arrays[cid])}
    print(cid, counts, int(arrays[cid].min()), int(arrays[cid].max()))

for cid, cfg, kind, stream, blocks in LONG_CASES:
    iq = make_input(cfg, kind, stream, blocks)
    # the 10 s capture is 30 720 000 bytes (192 k) / 38 400 000 (240 k): the tail short of a block is never demodulated (:863-868)
    tail = np.zeros(long_capture_bytes(cfg) - blocks * B, dtype=np.uint8)
    assert 0 <= tail.size < B
    pcm = RefOracle(**CONFIGS[cfg]).run(np.concatenate([iq, tail]))
    meta["long"][cid] = {"config": cfg, "kind": kind, "stream": stream, "blocks": blocks, "input_sha256": sha(iq),
                         "n_pcm": int(pcm.size), "pcm_sha256": sha(pcm)}
    print(cid, pcm.size)

for name, kw in [("192k_90", dict(rate_in=192000, size=90)), ("192k_128", dict(rate_in=192000, size=128)),
                 ("240k_90", dict(rate_in=240000, size=90)), ("240k_128", dict(rate_in=240000, size=128))]:
    t = RefOracle(mode=2, **kw).tables()
    for k, v in t.items():
        arrays[f"tab_{name}_{k}"] = v
    meta["tables"][name] = kw

r = RefOracle()
meta["layout"] = {str(i): r.layout(i) for i in range(27)}
np.savez_compressed(os.path.join(HERE, "golden.npz"), **arrays)
with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
    json.dump(meta, f, indent=1, sort_keys=True)
print("wrote", os.path.getsize(os.path.join(HERE, "golden.npz")), "bytes")
