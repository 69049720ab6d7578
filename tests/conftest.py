import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    # a machine without an NVIDIA device skips the gpu-marked tests (a GPU box where the library fails to load still FAILS them)
    if os.path.exists("/dev/nvidiactl") or os.path.exists("/dev/nvidia0"):
        return
    skip = pytest.mark.skip(reason="no NVIDIA device on this machine (agc_b200 has no CPU fallback)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


SYM2ASCII = np.full(64, ord('X'), np.uint8)
for _c, _s in zip("ACGTNRYSWKMBDHVU", range(16)):
    SYM2ASCII[_s] = ord(_c)
SYM2ASCII[30] = ord('X')
SYM2ASCII[32] = ord('@')


def to_fasta_body(codes, width=80, lower_frac=0.0, rng=None):
    """symbols -> raw FASTA body bytes (newline every `width` symbols), as CGenomeIO::ReadContigRaw returns them"""
    a = SYM2ASCII[np.asarray(codes, np.uint8)]
    if lower_frac and rng is not None and len(a):
        m = rng.random(len(a)) < lower_frac
        a = a.copy(); a[m] |= 0x20
    if width <= 0 or len(a) == 0:
        return a.tobytes() + (b"\n" if width > 0 else b"")
    rows = [a[i:i + width].tobytes() for i in range(0, len(a), width)]
    return b"\n".join(rows) + b"\n"


def mutate(rng, ref, p_snp=0.0, n_indel=0, n_nonacgt=0):
    t = np.array(ref, np.uint8, copy=True)
    if p_snp and len(t):
        m = rng.random(len(t)) < p_snp
        t[m] = (t[m] + rng.integers(1, 4, int(m.sum()))) % 4
    t = list(t)
    for _ in range(n_indel):
        pos = int(rng.integers(0, max(1, len(t))))
        if rng.random() < 0.5:
            del t[pos:pos + int(rng.integers(1, 30))]
        else:
            t[pos:pos] = list(rng.integers(0, 4, int(rng.integers(1, 30))))
    t = np.array(t, np.uint8)
    for _ in range(n_nonacgt):
        if len(t) == 0:
            break
        pos = int(rng.integers(0, len(t))); L = int(rng.integers(1, 12))
        t[pos:pos + L] = rng.choice([4, 4, 4, 5, 11, 30])
    return t


@pytest.fixture(scope="session")
def dev_factory():
    import agc_b200
    made = []

    def make(**kw):
        d = agc_b200.Device(**kw)
        made.append(d)
        return d
    yield make
    for d in made:
        d.close()
