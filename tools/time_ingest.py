"""Times the device-side TPED / .bed loaders against the host row packer on a configs[0]-sized cohort
(1 000 cases / 1 000 controls x 10 000 SNPs = 80 MB of TPED text). Usage: python tools/time_ingest.py [M] [N]"""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import libgwaspp_b200 as gw  # noqa: E402
from test_gpu_ingest import bed_encode, tped_bytes  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000
rng = np.random.default_rng(1)
codes = rng.choice(4, size=(M, N), p=[0.62, 0.3, 0.07, 0.01]).astype(np.uint8)
text = tped_bytes(codes)
d = tempfile.mkdtemp()
p = os.path.join(d, "c.tped")
open(p, "wb").write(text)
pb = os.path.join(d, "c.bed")
open(pb, "wb").write(bytes([0x6C, 0x1B, 0x01]) + bed_encode(codes).tobytes())
print(f"{M} SNPs x {N} samples: {len(text) / 1e6:.1f} MB of text, {os.path.getsize(pb) / 1e6:.2f} MB of .bed")


def best(f, n=5):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        f()
        ts.append(time.perf_counter() - t0)
    return min(ts)


t = best(lambda: gw.tped_dims(p))
print(f"tped_dims (host pass over the file)     {t * 1e3:8.2f} ms  {len(text) / t / 1e9:6.2f} GB/s")
with gw.GenoStore(M, N) as st:
    st.load_tped(p)
    t = best(lambda: st.load_tped(p))
    print(f"load_tped (file -> pinned -> device)    {t * 1e3:8.2f} ms  {len(text) / t / 1e9:6.2f} GB/s of text")
    rows_dev = st.get_rows()
    t = best(lambda: st.put_tped_text(text))
    print(f"put_tped_text (pageable host buffer)    {t * 1e3:8.2f} ms  {len(text) / t / 1e9:6.2f} GB/s of text")
    t = best(lambda: st.load_bed(pb))
    print(f"load_bed                                {t * 1e3:8.2f} ms  {os.path.getsize(pb) / t / 1e9:6.2f} GB/s of .bed")
    assert np.array_equal(st.get_rows(), rows_dev)
    # host packer: what the C++ mirror's addGenotypeRow loop costs (text already collapsed to 'XY\t')
    lut = np.array([np.frombuffer(s, np.uint8)[:3].tobytes() for s in (b"AA\t", b"AC\t", b"CC\t", b"00\t")])
    lines = [b"".join(lut[codes[r]].tolist()) for r in range(min(M, 500))]
    t0 = time.perf_counter()
    for l in lines:
        gw.pack_row_text(l, N)
    t = (time.perf_counter() - t0) * M / len(lines)
    print(f"host packer gwasdev_pack_row_text       {t * 1e3:8.2f} ms  (extrapolated from {len(lines)} rows, without reading or collapsing the text)")
