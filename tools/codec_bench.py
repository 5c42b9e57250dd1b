"""CPU benchmark of the PVM / DDS loader: the product's decoder (host/VolumeIO.cpp through
libvolren_host.so) beside the reference's own readPVMvolume (src/ddsbase.cpp compiled into oracle/_ref), both reading
the SAME .pvm file written by the reference encoder, payloads compared byte for byte.  Runs without a GPU.
usage: python tools/codec_bench.py [N=256] [reps=3]      (N^3 uint16 `mix` volume, C3's kind of data)"""
import os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "volume-renderer_b200", "python")]
from oracle import orc
from volren_b200 import host, workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dims = (n, n, n)
vol = orc.synth_mix(dims, 2, 4095, workloads.SEEDS["C3"], True, os.cpu_count() or 1)
with tempfile.TemporaryDirectory() as td:
    fn = os.path.join(td, "vol.pvm")
    t0 = time.perf_counter()
    orc.ref_write_pvm(fn, vol, dims, components=2)
    t_enc = time.perf_counter() - t0
    fbytes = os.path.getsize(fn)
    raw = vol.size * 2

    def best(f):
        ts, out = [], None
        for _ in range(reps):
            t = time.perf_counter(); out = f(); ts.append(time.perf_counter() - t)
        return min(ts), out

    t_ref, r = best(lambda: orc.ref_read_pvm(fn))
    t_our, o = best(lambda: host.pvm_decode(path=fn))
    data = open(fn, "rb").read()
    t_mem, m = best(lambda: host.pvm_decode(data=data))
    same = r is not None and o["ok"] and o["payload"] == r["payload"] == m["payload"] and tuple(o["dims"]) == tuple(r["dims"]) == dims
    print(f"{n}^3 uint16 mix volume: {raw / 2**20:.0f} MiB raw, {fbytes / 2**20:.1f} MiB as {data[:7].decode()} (reference encoder: {t_enc:.2f} s)")
    print(f"reference readPVMvolume (ddsbase.cpp, -O2):   {t_ref * 1e3:8.1f} ms = {raw / t_ref / 1e6:7.1f} MB/s of payload")
    print(f"product vrh_pvm_read (file):                  {t_our * 1e3:8.1f} ms = {raw / t_our / 1e6:7.1f} MB/s   ({t_ref / t_our:.2f}x)")
    print(f"product vrh_pvm_decode (bytes in memory):     {t_mem * 1e3:8.1f} ms = {raw / t_mem / 1e6:7.1f} MB/s   ({t_ref / t_mem:.2f}x)")
    print(f"payloads identical: {same}")
    sys.exit(0 if same else 1)
