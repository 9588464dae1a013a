"""Regenerates tests/golden/reference_cases.json from the oracle (run from the repo root:
python tests/golden/make_golden.py).  The JVM cannot run here, so these vectors are oracle
outputs whose TOTAL SIZES equal the sizes the reference itself logged."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import datasets  # noqa: E402
import orc  # noqa: E402

out = []
for data, name, chunk, expect in datasets.reference_cases():
    z = orc.compress(data, chunk, name, 0, literal=True)
    assert len(z) == expect, (name, len(z), expect)
    K = (len(data) + chunk - 1) // chunk
    lens = []
    for k in range(K):
        _, ln, _ = orc.encode_chunk(data[k * chunk:(k + 1) * chunk])
        lens.append("".join("%02x" % int(v) for v in ln))
    out.append({"name": name, "input_size": len(data), "input_sha256": hashlib.sha256(data).hexdigest(),
                "chunk_bytes": chunk, "dcz_size": len(z), "dcz_sha256": hashlib.sha256(z).hexdigest(),
                "payload_size": int.from_bytes(z[-8:], "big"), "code_lengths_hex": lens})
json.dump(out, open(os.path.join(HERE, "reference_cases.json"), "w"), indent=1)
print("wrote", len(out), "cases")
