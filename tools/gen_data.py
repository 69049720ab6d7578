"""Synthetic genome collections: NumPy PCG64, reference = iid uniform ACGT, sample = reference with iid
substitutions (+ optional indels / structural edits), FASTA upper-case, 80 columns, sample name = file stem."""
import os
import numpy as np

LET = np.frombuffer(b"ACGTN", np.uint8)


def write_fasta(path, contigs, width=80):
    """contigs: list of (name, symbol array 0..4)"""
    with open(path, "wb") as f:
        for name, codes in contigs:
            f.write(b">" + name.encode() + b"\n")
            a = LET[np.asarray(codes, np.uint8)]
            n = len(a)
            if n == 0:
                continue
            full = (n // width) * width
            if full:
                rows = a[:full].reshape(-1, width)
                out = np.empty((rows.shape[0], width + 1), np.uint8)
                out[:, :width] = rows; out[:, width] = 10
                f.write(out.tobytes())
            if n > full:
                f.write(a[full:].tobytes() + b"\n")


def substitute(rng, ref, p):
    t = ref.copy()
    if p > 0 and len(t):
        m = rng.random(len(t)) < p
        t[m] = (t[m] + rng.integers(1, 4, int(m.sum()), dtype=np.uint8)) % 4
    return t


def indels(rng, t, n, maxlen=50):
    t = list(t) if n else t
    for _ in range(n):
        pos = int(rng.integers(0, max(1, len(t))))
        L = int(rng.integers(1, maxlen + 1))
        if rng.random() < 0.5:
            del t[pos:pos + L]
        else:
            t[pos:pos] = list(rng.integers(0, 4, L))
    return np.asarray(t, np.uint8)


def viral(out_dir, n_samples=1000, ref_len=30000, p=0.01, seed=1):
    """config C2: one random reference, n_samples genomes with 1% SNPs, one contig each.  Returns file list (ref first)."""
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.default_rng(seed)
    ref = rng.integers(0, 4, ref_len, dtype=np.uint8)
    files = [os.path.join(out_dir, "ref.fa")]
    write_fasta(files[0], [("ref_ctg1", ref)])
    for i in range(n_samples):
        fn = os.path.join(out_dir, f"s{i:04d}.fa")
        write_fasta(fn, [(f"s{i:04d}_ctg1", substitute(rng, ref, p))])
        files.append(fn)
    return files, (n_samples + 1) * ref_len if p == 0 else None


def complex_collection(out_dir, seed=5, n_samples=12, ctg_len=40000, n_ctg=3, with_n=False):
    """small multi-contig collection that exercises the rare add_segment branches: deleted splitters (missing middle),
    one-sided segments, reverse-complemented contigs, novel contigs without splitters, duplicated sequences"""
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.default_rng(seed)
    ref = [rng.integers(0, 4, ctg_len + 1000 * i, dtype=np.uint8) for i in range(n_ctg)]
    files = [os.path.join(out_dir, "ref.fa")]
    write_fasta(files[0], [(f"chr{i + 1} some description", c) for i, c in enumerate(ref)])
    for s in range(n_samples):
        ctgs = []
        for i, c in enumerate(ref):
            t = substitute(rng, c, [0.0, 0.001, 0.01, 0.05][s % 4])
            t = indels(rng, t, s % 3)
            if s % 4 == 1 and len(t) > 9000:
                a = int(rng.integers(1000, len(t) - 8000)); t = np.concatenate([t[:a], t[a + int(rng.integers(1500, 6000)):]])   # big deletion
            if s % 5 == 2:
                t = (3 - t[::-1]).astype(np.uint8)                     # reverse complement
            if s % 6 == 3:
                t = t[int(rng.integers(100, 3000)):len(t) - int(rng.integers(100, 3000))]   # trimmed ends
            if with_n and s % 2 == 0 and len(t) > 5000:
                a = int(rng.integers(0, len(t) - 600)); t = t.copy(); t[a:a + int(rng.integers(1, 500))] = 4
            ctgs.append((f"smp{s}_chr{i + 1}", t))
        if s % 3 == 0:
            ctgs.append((f"smp{s}_novel", rng.integers(0, 4, int(rng.integers(50, 5000)), dtype=np.uint8)))
        if s % 4 == 2:
            ctgs.append((f"smp{s}_tiny", rng.integers(0, 4, 12, dtype=np.uint8)))
        if s % 7 == 5:
            ctgs = ctgs[::-1]
        if s == n_samples - 1:
            ctgs = [(f"smp{s}_copy{i}", c) for i, c in enumerate(ref)]   # identical to the reference: empty deltas
        fn = os.path.join(out_dir, f"smp{s:02d}.fa")
        write_fasta(fn, ctgs)
        files.append(fn)
    return files


def concatenated_collection(out_dir, seed=1, n_ctg=23, per_file=9):
    """-c mode: multi-genome FASTA files; every contig becomes a sample of its own (named after the contig) and the
    registration points come every pack_cardinality contigs, across file boundaries (agc_compressor.cpp:2180-2199)."""
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.default_rng(seed)
    ref = [rng.integers(0, 4, 20000 + 500 * i, dtype=np.uint8) for i in range(3)]
    files = [os.path.join(out_dir, "ref.fa")]
    write_fasta(files[0], [(f"chr{i + 1} desc", c) for i, c in enumerate(ref)])
    ctgs = []
    for j in range(n_ctg):
        t = indels(rng, substitute(rng, ref[j % 3], [0, 0.001, 0.02][j % 3]), j % 4)
        if j % 5 == 1:
            t = (3 - t[::-1]).astype(np.uint8)
        if j % 7 == 3:
            t = np.concatenate([t[:3000], t[9000:]])          # deleted splitters: no missing-middle split in -c mode (1366)
        if j % 11 == 4:
            t = rng.integers(0, 4, 3000, dtype=np.uint8)
        ctgs.append((f"g{j:03d} genome {j}", t))
    for f in range(0, n_ctg, per_file):
        fn = os.path.join(out_dir, f"part{f:03d}.fa")
        write_fasta(fn, ctgs[f:f + per_file])
        files.append(fn)
    return files


def fallback_collection(out_dir, seed=11, seg=2000, n_samples=8, ref_len=60000):
    """-f mode: divergent samples (up to 6 % substitutions: most splitters are destroyed, so segments come with one or no
    terminal splitter, or with a pair that is not in the map) plus reference fragments that contain no whole splitter pair:
    the cases find_cand_segment_using_fallback_minimizers decides (agc_compressor.cpp:1290,1327,1352,1461)."""
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.default_rng(seed)
    ref = [rng.integers(0, 4, ref_len // 2 + 901 * i, dtype=np.uint8) for i in range(2)]
    files = [os.path.join(out_dir, "ref.fa")]
    write_fasta(files[0], [(f"chr{i + 1}", c) for i, c in enumerate(ref)])
    for s in range(n_samples):
        ctgs = []
        for i, c in enumerate(ref):
            t = indels(rng, substitute(rng, c, [0.002, 0.02, 0.06][s % 3]), 4)
            if s % 4 == 1:
                t = (3 - t[::-1]).astype(np.uint8)
            ctgs.append((f"f{s}_chr{i + 1}", t))
        a = int(rng.integers(0, len(ref[0]) - 3 * seg))
        frag = substitute(rng, ref[0][a:a + int(2.2 * seg)], 0.03)
        ctgs.append((f"f{s}_frag", frag if s % 2 else (3 - frag[::-1]).astype(np.uint8)))
        ctgs.append((f"f{s}_novel", rng.integers(0, 4, int(1.5 * seg), dtype=np.uint8)))
        fn = os.path.join(out_dir, f"f{s:02d}.fa")
        write_fasta(fn, ctgs)
        files.append(fn)
    return files


def total_bases(files):
    n = 0
    for fn in files:
        with open(fn, "rb") as f:
            for line in f:
                if not line.startswith(b">"):
                    n += len(line.rstrip(b"\r\n"))
    return n


def adaptive_collection(out_dir, seed=2, n_samples=8, ref_len=60000, n_ctg=2, novel_len=9000, p=0.01):
    """config C3 in small: every sample = mutated reference contigs + 20 random 1-50 b indels + one NOVEL random contig
    (>= segment_size, so `-a` looks for new splitters in it, agc_compressor.cpp:2038-2044).  Later samples also carry a
    mutated copy of an earlier sample's novel contig (cut by the splitters that sample added), a short novel contig
    (< segment_size: set aside but not searched) and, once, a contig that only occurs reverse-complemented."""
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.default_rng(seed)
    ref = [rng.integers(0, 4, ref_len // n_ctg + 777 * i, dtype=np.uint8) for i in range(n_ctg)]
    files = [os.path.join(out_dir, "ref.fa")]
    write_fasta(files[0], [(f"chr{i + 1}", c) for i, c in enumerate(ref)])
    novel = []
    for s in range(n_samples):
        ctgs = [(f"a{s}_chr{i + 1}", indels(rng, substitute(rng, c, p), 20 // n_ctg)) for i, c in enumerate(ref)]
        nv = rng.integers(0, 4, novel_len + 531 * s, dtype=np.uint8)
        ctgs.append((f"a{s}_novel", nv))
        if novel:
            old = novel[int(rng.integers(0, len(novel)))]
            t = indels(rng, substitute(rng, old, p), 3)
            if s % 3 == 2:
                t = (3 - t[::-1]).astype(np.uint8)
            ctgs.append((f"a{s}_seen_before", t))
        if s % 2 == 1:
            ctgs.append((f"a{s}_short_novel", rng.integers(0, 4, 700, dtype=np.uint8)))
        if s == 4:
            ctgs.append((f"a{s}_twice", nv[::-1].copy()))     # shares no canonical k-mer with nv: a second searched contig
        novel.append(nv)
        fn = os.path.join(out_dir, f"a{s:02d}.fa")
        write_fasta(fn, ctgs)
        files.append(fn)
    return files


def bacterial_adaptive(out_dir, seed=2, n_samples=63, ref_len=5_000_000, p=0.01, novel_len=120_000, n_indels=20):
    """BASELINE configs[2] / SURVEY C3: ref 5 Mb; every sample = the reference with 1 % substitutions and 20 random 1-50 b indels
    plus one NOVEL random 120 kb contig (>= segment_size: `-a` finds new splitters in it).  Returns the file list (ref first)."""
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.default_rng(seed)
    ref = rng.integers(0, 4, ref_len, dtype=np.uint8)
    files = [os.path.join(out_dir, "ref.fa")]
    write_fasta(files[0], [("ref_chr", ref)])
    for s in range(n_samples):
        t = substitute(rng, ref, p)
        pos = np.sort(rng.integers(0, len(t), n_indels))[::-1]
        for q in pos:                                     # few large-array edits instead of list surgery
            L = int(rng.integers(1, 51))
            if rng.random() < 0.5:
                t = np.concatenate([t[:q], t[q + L:]])
            else:
                t = np.concatenate([t[:q], rng.integers(0, 4, L, dtype=np.uint8), t[q:]])
        fn = os.path.join(out_dir, f"b{s:02d}.fa")
        write_fasta(fn, [(f"b{s:02d}_chr", t), (f"b{s:02d}_novel", rng.integers(0, 4, novel_len, dtype=np.uint8))])
        files.append(fn)
    return files


def human_chromosome(out_dir, seed=3, n_samples=1, ctg_len=250_000_000, n_repeats=200, p=0.001, indel_every=10_000):
    """BASELINE configs[3] / SURVEY C4 in shape: the reference is ONE long contig with repeat content (n_repeats random 1-10 kb
    blocks copied to random places); every sample = the reference with substitutions at rate p and one random 1-50 b indel per
    indel_every bases.  Returns the file list (ref first)."""
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.default_rng(seed)
    ref = rng.integers(0, 4, ctg_len, dtype=np.uint8)
    for _ in range(n_repeats):
        L = int(rng.integers(1000, 10001))
        a = int(rng.integers(0, ctg_len - L)); b = int(rng.integers(0, ctg_len - L))
        ref[b:b + L] = ref[a:a + L]
    files = [os.path.join(out_dir, "ref.fa")]
    write_fasta(files[0], [("chr1", ref)])
    for s in range(n_samples):
        t = substitute(rng, ref, p)
        n_ind = max(1, ctg_len // indel_every)
        pos = np.sort(rng.integers(0, len(t), n_ind))
        lens = rng.integers(1, 51, n_ind)
        dele = rng.random(n_ind) < 0.5
        pieces, prev = [], 0
        for q, L, d in zip(pos, lens, dele):                # one pass: pieces between the edit points
            q = int(q)
            if q < prev:
                continue
            pieces.append(t[prev:q])
            if d:
                prev = min(len(t), q + int(L))
            else:
                pieces.append(rng.integers(0, 4, int(L), dtype=np.uint8)); prev = q
        pieces.append(t[prev:])
        t = np.concatenate(pieces)
        fn = os.path.join(out_dir, f"h{s:02d}.fa")
        write_fasta(fn, [(f"h{s:02d}_chr1", t)])
        files.append(fn)
    return files
