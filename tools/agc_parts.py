"""Structural reader of .agc containers (src/common/archive.cpp:142-169,280-293) and of agc-b200's --dump-parts files.
TEST / DIAGNOSTIC TOOL: uses the reference's own libzstd (oracle/_ref/libzstd_ref.so) to look inside parts."""
import ctypes as C
import os
import struct

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_Z = None


def zstd_ref():
    global _Z
    if _Z is None:
        _Z = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libzstd_ref.so"))
        _Z.ZSTD_decompress.restype = C.c_size_t
        _Z.ZSTD_decompress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        _Z.ZSTD_compress.restype = C.c_size_t
        _Z.ZSTD_compress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int]
        _Z.ZSTD_compressBound.restype = C.c_size_t
        _Z.ZSTD_compressBound.argtypes = [C.c_size_t]
        _Z.ZSTD_getFrameContentSize.restype = C.c_ulonglong
        _Z.ZSTD_getFrameContentSize.argtypes = [C.c_char_p, C.c_size_t]
        _Z.ZSTD_isError.restype = C.c_uint
        _Z.ZSTD_isError.argtypes = [C.c_size_t]
    return _Z


def zstd_decompress(frame: bytes) -> bytes:
    z = zstd_ref()
    n = z.ZSTD_getFrameContentSize(frame, len(frame))
    assert n < (1 << 40), "bad frame"
    out = C.create_string_buffer(max(int(n), 1))
    r = z.ZSTD_decompress(out, int(n), frame, len(frame))
    assert not z.ZSTD_isError(r), "zstd decode error"
    return out.raw[:r]


def zstd_compress(raw: bytes, level: int) -> bytes:
    """ZSTD_compress of the vendored 1.5.5 == ZSTD_compressCCtx on a fresh context (frames do not depend on ctx history)"""
    z = zstd_ref()
    cap = z.ZSTD_compressBound(len(raw))
    out = C.create_string_buffer(cap)
    r = z.ZSTD_compress(out, cap, raw, len(raw), level)
    assert not z.ZSTD_isError(r)
    return out.raw[:r]


def _varint(b, p):
    n = b[p]; p += 1
    v = 0
    for _ in range(n):
        v = (v << 8) | b[p]; p += 1
    return v, p


def read_archive(path):
    """-> (stream_names, parts) ; parts = list of dict(stream, name, index, metadata, payload) in file order"""
    b = open(path, "rb").read()
    fs = struct.unpack("<Q", b[-8:])[0]
    p = len(b) - 8 - fs
    ns, p = _varint(b, p)
    names, parts = [], []
    for s in range(ns):
        q = b.index(b"\0", p)
        name = b[p:q].decode("latin1"); p = q + 1
        np_, p = _varint(b, p)
        _raw, p = _varint(b, p)
        names.append(name)
        for i in range(np_):
            off, p = _varint(b, p)
            size, p = _varint(b, p)
            meta, dp = _varint(b, off)
            parts.append(dict(stream=s, name=name, index=i, metadata=meta, payload=b[dp:dp + size], offset=off))
    parts.sort(key=lambda x: x["offset"])
    return names, parts


def _cvar(b, p):
    """collection.h:162-196 prefix varint"""
    c = b[p]
    if c < 0x80: return c, p + 1
    if c < 0xC0: return ((c << 8) + b[p + 1]) + (1 << 7) - (0x80 << 8), p + 2
    if c < 0xE0: return ((c << 16) + (b[p + 1] << 8) + b[p + 2]) + (1 << 7) + (1 << 14) - (0xC0 << 16), p + 3
    if c < 0xF0: return ((c << 24) + (b[p + 1] << 16) + (b[p + 2] << 8) + b[p + 3]) + (1 << 7) + (1 << 14) + (1 << 21) - (0xE0 << 24), p + 4
    v = (b[p + 1] << 24) + (b[p + 2] << 16) + (b[p + 3] << 8) + b[p + 4]
    return v + (1 << 7) + (1 << 14) + (1 << 21) + (1 << 28), p + 5


def part_contents(part):
    """decode one archive part into the list of pre-zstd byte strings it was built from, plus a 'form' tag"""
    name, meta, pl = part["name"], part["metadata"], part["payload"]
    if name in ("params", "splitters", "segment-splitters", "file_type_info"):
        return "immediate", [pl]
    if name in ("collection-samples", "collection-contigs"):
        return "frame", [zstd_decompress(pl)]
    if name == "collection-details":
        p = 0; sizes = []
        for _ in range(5):
            r, p = _cvar(pl, p); k, p = _cvar(pl, p); sizes.append((r, k))
        out = []
        for r, k in sizes:
            d = zstd_decompress(pl[p:p + k]); assert len(d) == r; out.append(d); p += k
        return "details", out
    if meta == 0:
        return "raw", [pl]
    return ("tuples" if pl[-1] == 1 else "plain"), [zstd_decompress(pl[:-1])]


def read_dump(path):
    """agc-b200 --dump-parts file -> (stream_names, records)"""
    b = open(path, "rb").read()
    p = 0
    recs, names = [], []

    def u64():
        nonlocal p
        v = struct.unpack_from("<Q", b, p)[0]; p += 8
        return v

    def blob():
        nonlocal p
        n = u64(); v = b[p:p + n]; p += n
        return v
    while p < len(b):
        tag = b[p:p + 4]; p += 4
        if tag == b"PART":
            sid = u64(); ep = u64(); kind = u64(); raw_size = u64(); nt = u64()
            tasks = []
            for _ in range(nt):
                lvl = u64(); tasks.append((lvl, blob()))
            fb = blob()
            recs.append(dict(kind=int(kind), stream=int(sid), epoch=ep, raw_size=raw_size, tasks=tasks, fallback=fb))
        elif tag == b"IMMD":
            nm = blob().decode(); meta = u64(); d = blob()
            recs.append(dict(kind=9, name=nm, metadata=meta, tasks=[(0, d)]))
        elif tag == b"STRM":
            names.append(blob().decode("latin1"))
        else:
            raise ValueError("bad dump tag %r at %d" % (tag, p - 4))
    return names, recs


def compare_dump_to_archive(dump_path, agc_path):
    """returns list of mismatch descriptions (empty = the two writers produced the same parts in the same order)"""
    dn, recs = read_dump(dump_path)
    an, parts = read_archive(agc_path)
    bad = []
    if dn != an:
        bad.append(f"stream directory differs: {len(dn)} vs {len(an)} streams; first diff at "
                   f"{next((i for i, (x, y) in enumerate(zip(dn, an)) if x != y), min(len(dn), len(an)))}")
    if len(recs) != len(parts):
        bad.append(f"part count differs: dump {len(recs)} vs archive {len(parts)}")
    for i, (r, a) in enumerate(zip(recs, parts)):
        form, content = part_contents(a)
        if r["kind"] == 9:
            if a["name"] != r["name"] or a["metadata"] != r["metadata"] or content[0] != r["tasks"][0][1]:
                bad.append(f"part {i}: immediate part {r['name']} differs from archive part {a['name']}")
            continue
        nm = dn[r["stream"]] if r["stream"] < len(dn) else "?"
        if nm != a["name"]:
            bad.append(f"part {i}: stream {nm} vs archive {a['name']}[{a['index']}]")
            continue
        if form == "raw":
            if r["fallback"] != content[0]:
                bad.append(f"part {i} ({nm}): raw-stored part differs")
            continue
        exp_kind = {"plain": 0, "tuples": 1, "frame": 2, "details": 3}[form]
        if r["kind"] != exp_kind:
            bad.append(f"part {i} ({nm}): kind {r['kind']} vs archive form {form}")
            continue
        if form in ("plain", "tuples", "frame") and a["metadata"] != r["raw_size"]:
            bad.append(f"part {i} ({nm}): metadata {r['raw_size']} vs {a['metadata']}")
        got = [t[1] for t in r["tasks"]]
        if got != content:
            k = next(j for j, (x, y) in enumerate(zip(got, content)) if x != y) if len(got) == len(content) else -1
            bad.append(f"part {i} ({nm}[{a['index']}]): content differs (task {k}, {len(got[k]) if k >= 0 else '?'} vs {len(content[k]) if k >= 0 else '?'} bytes)")
        if len(bad) > 20:
            break
    return bad


if __name__ == "__main__":
    import sys
    if len(sys.argv) == 3:
        for m in compare_dump_to_archive(sys.argv[1], sys.argv[2]) or ["identical part sequence"]:
            print(m)
    else:
        names, parts = read_archive(sys.argv[1])
        for pt in parts:
            form, c = part_contents(pt)
            print(f"@{pt['offset']:>10} {pt['name']:<22}[{pt['index']}] meta={pt['metadata']:<8} {len(pt['payload']):>8} B  {form} {[len(x) for x in c]}")
