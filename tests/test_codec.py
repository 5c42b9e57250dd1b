"""CPU: the PVM/DDS codec.  Three decoders must agree byte for byte on every input:
  (1) the reference's own src/ddsbase.cpp compiled unmodified into oracle/_ref (when present),
  (2) the oracle restatement oracle/codec_oracle.c,
  (3) the product decoder volume-renderer_b200/host/VolumeIO.cpp.
Golden .pvm fixtures were written by the reference encoder (tests/golden/make_pvm_fixtures.py)
and carry the reference's own checksum() of their payload."""
import json
import os

import numpy as np
import pytest

from oracle import orc
from volren_b200 import host, workloads

FIX = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pvm_fixtures.json")))
have_ref = orc.ref_lib() is not None


@pytest.mark.parametrize("name", sorted(FIX))
def test_golden_fixture_decodes_identically(name, golden_dir):
    meta = FIX[name]
    path = os.path.join(golden_dir, name)
    raw = open(path, "rb").read()
    assert len(raw) == meta["file_bytes"]
    o = orc.pvm_decode(raw)
    p = host.pvm_decode(data=raw)
    p2 = host.pvm_decode(path=path)
    assert o is not None and p["ok"] and p2["ok"]
    n = meta["payload_bytes"]
    assert o["dims"] == tuple(meta["dims"]) == p["dims"] == p2["dims"]
    assert o["components"] == meta["components"] == p["components"]
    assert o["scale"] == pytest.approx(meta["scale"]) and p["scale"] == pytest.approx(meta["scale"])
    assert o["payload"][:n] == p["payload"] == p2["payload"]
    # the reference's own digest (ddsbase.cpp:872-893) of the payload, recorded at fixture time
    assert orc.checksum(p["payload"]) == meta["ref_checksum"] == host.dds_checksum(p["payload"])
    if meta["strings"]:
        s = meta["strings"]
        assert p["strings"] == [s["description"], s["courtesy"], s["parameter"], s["comment"]]
        assert p["version"] == 3
    if have_ref:
        r = orc.ref_read_pvm(path)
        assert r["payload"] == p["payload"] and r["dims"] == p["dims"] and r["components"] == p["components"]
        assert orc.ref_checksum(r["payload"]) == meta["ref_checksum"]


def test_fixture_payloads_are_the_seeded_volumes(golden_dir):
    v = workloads.mix_volume((24, 20, 16), 255, 0xA1, with_hash=False)
    assert host.pvm_decode(path=os.path.join(golden_dir, "pvm1_u8_24x20x16.pvm"))["payload"] == v.tobytes()
    v = workloads.mix_volume((20, 18, 12), 4095, 0xA2, with_hash=True)
    assert host.pvm_decode(path=os.path.join(golden_dir, "pvm2_u16_20x18x12.pvm"))["payload"] == v.tobytes()


@pytest.mark.skipif(not have_ref, reason="oracle/_ref (reference ddsbase.cpp) not built")
@pytest.mark.parametrize("dims,comps,scale,kind", [
    ((33, 17, 9), 1, (1.0, 1.0, 1.0), "smooth"),
    ((16, 16, 16), 2, (1.0, 1.0, 1.0), "smooth"),
    ((7, 5, 3), 1, (2.0, 1.0, 0.5), "noise"),
    ((1, 1, 1), 1, (1.0, 1.0, 1.0), "noise"),
    ((64, 1, 1), 2, (1.0, 1.0, 3.0), "noise"),
    ((5, 4, 3), 3, (1.0, 1.0, 1.0), "noise"),          # 3 interleaved channels
    ((40, 30, 2), 4, (1.0, 1.0, 1.0), "zeros"),
    ((31, 29, 7), 1, (1.0, 1.0, 1.0), "ramp"),
])
def test_round_trip_through_the_reference_encoder(tmp_path, dims, comps, scale, kind):
    n = dims[0] * dims[1] * dims[2] * comps
    rng = np.random.default_rng(n)
    if kind == "smooth":
        base = workloads.mix_volume(dims, 255 if comps == 1 else 4095, 3, with_hash=False)
        vol = base.view(np.uint8)[:n] if base.dtype == np.uint16 else base
        vol = np.ascontiguousarray(vol)
    elif kind == "noise":
        vol = rng.integers(0, 256, n, dtype=np.uint8)
    elif kind == "zeros":
        vol = np.zeros(n, np.uint8)
    else:
        vol = (np.arange(n) % 251).astype(np.uint8)
    path = str(tmp_path / "t.pvm")
    orc.ref_write_pvm(path, vol, dims, comps, scale)
    raw = open(path, "rb").read()
    r = orc.ref_read_pvm(path)
    o = orc.pvm_decode(raw)
    p = host.pvm_decode(data=raw)
    assert r["payload"] == vol.tobytes()
    assert o["payload"] == vol.tobytes()
    assert p["ok"] and p["payload"] == vol.tobytes()
    assert p["dims"] == tuple(dims) and p["components"] == comps and p["scale"] == pytest.approx(scale)


@pytest.mark.skipif(not have_ref, reason="oracle/_ref (reference ddsbase.cpp) not built")
def test_v3e_block_interleave_above_16MiB(tmp_path):
    """Payloads > 2^24 bytes switch the container to "DDS v3e" with per-block interleave
    (ddsbase.cpp:531,589).  272x256x128 x 2 components = 17 MiB + header."""
    dims = (272, 256, 128)
    vol = workloads.mix_volume(dims, 4095, 17, with_hash=False)
    path = str(tmp_path / "big.pvm")
    orc.ref_write_pvm(path, vol, dims, 2, (1.0, 1.0, 1.0))
    raw = open(path, "rb").read()
    assert raw[:8] == b"DDS v3e\n"
    p = host.pvm_decode(data=raw)
    assert p["ok"] and p["payload"] == vol.tobytes()
    o = orc.pvm_decode(raw)
    assert o["payload"] == vol.tobytes()
    assert host.dds_checksum(p["payload"]) == orc.ref_checksum(vol.tobytes())


def test_malformed_inputs_are_rejected_not_crashed(golden_dir):
    good = open(os.path.join(golden_dir, "pvm1_u8_24x20x16.pvm"), "rb").read()
    for bad in (b"", b"PVM", b"DDS v3d\n", b"XYZ\n1 1 1\n1\n\0", b"PVM\n0 1 1\n1\n", b"PVM\n2 2 2\n1\nabc",
                good[:len(good) // 2], b"PVM2\n2 2 2\n1 1 x\n1\n" + bytes(8), b"PVM\n2 2 2\n0\n" + bytes(8)):
        p = host.pvm_decode(data=bad)
        assert not p["ok"] and p["error"]
        assert orc.pvm_decode(bad) is None
    # trailing garbage after a complete payload
    assert not host.pvm_decode(data=b"PVM\n1 1 1\n1\nAB")["ok"]


def test_plain_pvm_with_comment_lines():
    raw = b"PVM\n# a comment\n# another\n2 2 1\n1\n" + bytes([1, 2, 3, 4])
    p = host.pvm_decode(data=raw)
    o = orc.pvm_decode(raw)
    assert p["ok"] and p["dims"] == (2, 2, 1) and p["payload"] == bytes([1, 2, 3, 4]) == o["payload"]


def test_truncated_and_corrupted_streams_never_crash_and_agree_with_the_oracle(golden_dir):
    """Fuzz around a valid DDS file: every truncation point near the head and the tail plus seeded byte
    flips.  The product decoder (64-bit bit reader, unchecked field reads inside a run) and the oracle
    decoder (bit-at-a-time restatement) must agree on every byte whenever the product decoder accepts."""
    good = open(os.path.join(golden_dir, "pvm1_u8_24x20x16.pvm"), "rb").read()
    assert good[:8] in (b"DDS v3d\n", b"DDS v3e\n")
    rng = np.random.default_rng(99)
    variants = [good[:k] for k in list(range(8, 40)) + list(range(len(good) - 24, len(good)))]
    for _ in range(60):
        b = bytearray(good)
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(8, len(b)))] ^= 1 << int(rng.integers(0, 8))
        variants.append(bytes(b))
    for v in variants:
        mine, ref = host.pvm_decode(data=v), orc.pvm_decode(v)
        if mine["ok"]:
            assert ref is not None
            assert mine["dims"] == tuple(ref["dims"]) and bytes(mine["payload"]) == bytes(ref["payload"])
        elif ref is not None:
            # documented deviation: the reference decodes zero padding when the end
            # marker is missing ("print and continue"); the product decoder reports the truncation
            assert "end-of-stream" in mine["error"]
