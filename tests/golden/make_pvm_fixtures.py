"""Writes the .pvm golden fixtures with the REFERENCE's own encoder (oracle/_ref, compiled from
/root/reference/src/ddsbase.cpp) and records the reference's own checksum() of each payload.
Needs /root/reference (build container only); the fixtures themselves are committed.

    python tests/golden/make_pvm_fixtures.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "volume-renderer_b200", "python"), os.path.dirname(HERE)]

from oracle import orc                      # noqa: E402
from volren_b200 import workloads           # noqa: E402


def main():
    assert orc.ref_lib() is not None, "oracle/_ref not built (needs /root/reference)"
    meta = {}

    def emit(name, vol, dims, comps, scale, **strings):
        path = os.path.join(HERE, name)
        orc.ref_write_pvm(path, vol, dims, comps, scale, **strings)
        back = orc.ref_read_pvm(path)
        assert back["payload"] == vol.tobytes()
        meta[name] = {"dims": list(dims), "components": comps, "scale": list(scale),
                      "payload_bytes": len(back["payload"]), "ref_checksum": orc.ref_checksum(back["payload"]),
                      "file_bytes": os.path.getsize(path), "strings": strings}
        print(name, meta[name])

    v = workloads.mix_volume((24, 20, 16), 255, 0xA1, with_hash=False)
    emit("pvm1_u8_24x20x16.pvm", v, (24, 20, 16), 1, (1.0, 1.0, 1.0))
    v = workloads.mix_volume((20, 18, 12), 4095, 0xA2, with_hash=True)
    emit("pvm2_u16_20x18x12.pvm", v, (20, 18, 12), 2, (1.0, 1.5, 2.0))
    rng = np.random.default_rng(5)
    v = rng.integers(0, 256, 9 * 7 * 5, dtype=np.uint8)
    emit("pvm3_u8_9x7x5_strings.pvm", v, (9, 7, 5), 1, (0.5, 0.25, 1.0),
         description="synthetic noise", courtesy="volren_b200 tests", parameter="9x7x5", comment="golden")
    # plain (not DDS-wrapped) PVM: readPVMvolume falls back to readRAWfile (ddsbase.cpp:783-784)
    v = np.arange(4 * 3 * 2, dtype=np.uint8)
    with open(os.path.join(HERE, "pvm_plain_4x3x2.pvm"), "wb") as f:
        f.write(b"PVM\n4 3 2\n1\n" + v.tobytes())
    back = orc.ref_read_pvm(os.path.join(HERE, "pvm_plain_4x3x2.pvm"))
    assert back["payload"] == v.tobytes()
    meta["pvm_plain_4x3x2.pvm"] = {"dims": [4, 3, 2], "components": 1, "scale": [1.0, 1.0, 1.0], "payload_bytes": 24,
                                   "ref_checksum": orc.ref_checksum(v.tobytes()), "file_bytes": 36, "strings": {}}
    with open(os.path.join(HERE, "pvm_fixtures.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
