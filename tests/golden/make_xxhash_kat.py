"""Generates tests/golden/xxhash_kat.json from python-xxhash (the package the reference hashes with:
pyproject.toml `xxhash`; xxh3_64_intdigest at probabilistic_single_filter_model.py:88,155-158; XXH64 with
seed j is what cobs hashes with).  Run here (python-xxhash 3.7.0 / libxxhash 0.8.2); the JSON is committed
so the oracle and the device functions stay pinned on boxes without python-xxhash."""
import json
import random
from pathlib import Path

import xxhash

rng = random.Random(20261018)
vecs = []
fixed = [b"", b"A", b"AC", b"ACG", b"ACGT", b"AGAGATTACGTCTGGTTGCAA", b"TAAATAAATTTATATAGCTAA", b"AAATAAATTTATATAGCTAAA"]
for n in list(range(0, 41)) + [48, 63, 64, 65, 96, 127, 128]:
    for _ in range(3):
        fixed.append(bytes(rng.choice(b"ACGTNacgtRYKM") for _ in range(n)))
for data in fixed:
    vecs.append({
        "data": data.decode("ascii"),
        "xxh64": {str(s): xxhash.xxh64_intdigest(data, seed=s) for s in (0, 1, 2, 3, 4, 5, 6, 7, 31, 63)},
        "xxh3_64": xxhash.xxh3_64_intdigest(data),
    })
out = {"generator": f"python-xxhash {xxhash.VERSION} / libxxhash {xxhash.XXHASH_VERSION}", "vectors": vecs}
Path(__file__).with_name("xxhash_kat.json").write_text(json.dumps(out, indent=0))
print(len(vecs), "vectors")
