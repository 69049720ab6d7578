"""GPU parity: every kernel behind the C ABI against the C oracle (oracle/agc_oracle.c) on the same seeded inputs.
Bit-exact (integer/byte work)."""
import numpy as np
import pytest
import orc
from conftest import to_fasta_body, mutate

pytestmark = pytest.mark.gpu


def _mk_pairs(rng, n_pairs, dirty=False):
    """(ref, text) pairs covering equal / SNP / indel / unrelated / tiny cases"""
    pairs = []
    for it in range(n_pairs):
        m = int(rng.choice([0, 7, 40, 300, 2000, 9000, 30000]))
        ref = rng.integers(0, 4, m).astype(np.uint8)
        if it % 7 == 3 and m > 200:
            ref[100:180] = np.tile(ref[100:105], 16)            # short tandem repeat -> long probe chains
        if dirty and it % 2 == 0:
            ref = mutate(rng, ref, 0, 0, 2)
        kind = it % 6
        if kind == 0:
            text = ref.copy()
        elif kind == 1:
            text = mutate(rng, ref, 0.001, 0)
        elif kind == 2:
            text = mutate(rng, ref, 0.02, 2)
        elif kind == 3:
            text = mutate(rng, ref, 0.2, 1)
        elif kind == 4:
            text = rng.integers(0, 4, int(rng.integers(0, 500))).astype(np.uint8)
        else:
            text = mutate(rng, ref, 0.005, 3)
        if dirty:
            text = mutate(rng, text, 0, 0, int(rng.integers(1, 4)))
            if it % 5 == 0 and len(text) > 60:
                text[20:20 + int(rng.integers(3, 40))] = 4           # N run
        pairs.append((ref, text))
    return pairs


def _run_pairs(dev, rng, pairs, mml, use_rc, check_decode=False):
    """upload refs and texts as contigs, build references from segments, encode/estimate/cost-vector on device"""
    contigs, reqs_ref, reqs_txt, expect_text = [], [], [], []
    for g, (ref, text) in enumerate(pairs):
        rc_ref = use_rc and g % 3 == 1
        rc_txt = use_rc and g % 2 == 1
        # what is uploaded is the reverse complement when the request asks for is_rc (so the kernel sees ref/text)
        up_ref = orc.revcomp(ref) if rc_ref else ref
        up_txt = orc.revcomp(text) if rc_txt else text
        pad_l = rng.integers(0, 4, int(rng.integers(0, 70))).astype(np.uint8)      # segments start at odd base offsets
        pad_r = rng.integers(0, 4, int(rng.integers(0, 70))).astype(np.uint8)
        contigs.append(to_fasta_body(np.concatenate([pad_l, up_ref, pad_r]), 80, 0.3, rng))
        reqs_ref.append((2 * g, len(pad_l), len(ref), rc_ref, 16 + g))
        contigs.append(to_fasta_body(np.concatenate([pad_r, up_txt, pad_l]), 61))
        reqs_txt.append((2 * g + 1, len(pad_r), len(text), rc_txt, 16 + g))
        expect_text.append(text)
    dev.set_splitters(np.zeros(0, np.uint64))
    dev.scan_contigs(contigs)
    for (c, s, l, rc, g), (ref, text) in zip(reqs_ref, pairs):
        assert np.array_equal(dev.get_segment(c, s, l, rc), ref)
    for (c, s, l, rc, g), text in zip(reqs_txt, expect_text):
        assert np.array_equal(dev.get_segment(c, s, l, rc), text)
    dev.put_references(reqs_ref)
    enc = dev.lz_encode(reqs_txt)
    est = dev.lz_estimate(reqs_txt)
    bounds = [int(rng.integers(0, 60)) for _ in reqs_txt]
    est_b = dev.lz_estimate([r + (b,) for r, b in zip(reqs_txt, bounds)])
    for g, (ref, text) in enumerate(pairs):
        z = orc.LZ(ref, mml)
        ht, short = z.ht()
        assert np.array_equal(dev.get_index(16 + g), ht), f"index layout differs, pair {g}"
        assert enc[g] == z.encode(text), f"encode differs, pair {g} (m={len(ref)}, n={len(text)})"
        assert int(est[g]) == z.estimate(text), f"estimate differs, pair {g}"
        assert int(est_b[g]) == z.estimate(text, bounds[g]), f"bounded estimate differs, pair {g}"
        if g % 4 == 0 and len(text):
            for pf in (0, 1):
                cv = dev.lz_cost_vector(reqs_txt[g], pf)
                assert np.array_equal(cv, z.cost_vector(text, pf)), f"cost vector differs, pair {g} prefix={pf}"
        if len(text) and text.max() <= 20 and len(enc[g]):
            assert np.array_equal(orc.lz_decode(ref, enc[g], mml, len(text)), text)
    if check_decode:                         # CLZDiff_V2::Decode on the device: every delta gives back its text
        ok = [g for g, (ref, text) in enumerate(pairs) if len(text) == 0 or text.max() <= 20]     # literals above 'A'+20 have no decoding (lz_diff.h is_literal)
        dec = dev.lz_decode([16 + g for g in ok], [enc[g] for g in ok])
        for g, d in zip(ok, dec):
            if len(enc[g]) == 0:
                assert len(d) == 0           # "equal to the reference" is stored as an empty delta
            else:
                assert np.array_equal(d, pairs[g][1]), f"device decode differs, pair {g}"


@pytest.mark.parametrize("mml", [15, 20, 32])
def test_lz_clean(dev_factory, mml):
    rng = np.random.default_rng(100 + mml)
    dev = dev_factory(k=21, min_match_len=mml)
    _run_pairs(dev, rng, _mk_pairs(rng, 36), mml, use_rc=False)


def test_lz_reverse_complement(dev_factory):
    rng = np.random.default_rng(7)
    dev = dev_factory(k=21, min_match_len=20)
    _run_pairs(dev, rng, _mk_pairs(rng, 30), 20, use_rc=True)


def test_lz_non_acgt(dev_factory):
    rng = np.random.default_rng(8)
    dev = dev_factory(k=21, min_match_len=18)
    _run_pairs(dev, rng, _mk_pairs(rng, 30, dirty=True), 18, use_rc=True)


def test_lz_many_segments_one_group(dev_factory):
    """many texts against one reference: exercises the shared-memory staged path and unit splitting"""
    rng = np.random.default_rng(9)
    mml = 20
    dev = dev_factory(k=25, min_match_len=mml)
    ref = rng.integers(0, 4, 30000).astype(np.uint8)
    texts = [mutate(rng, ref, 0.01, int(rng.integers(0, 3))) for _ in range(150)]
    contigs = [to_fasta_body(ref)] + [to_fasta_body(t) for t in texts]
    dev.set_splitters(np.zeros(0, np.uint64))
    dev.scan_contigs(contigs)
    dev.put_references([(0, 0, len(ref), False, 16)])
    reqs = [(i + 1, 0, len(t), False, 16) for i, t in enumerate(texts)]
    enc = dev.lz_encode(reqs)
    z = orc.LZ(ref, mml)
    for i, t in enumerate(texts):
        assert enc[i] == z.encode(t), i
    st = dev.stats()
    assert st.lz_alg_bytes > 0 and st.last_lz_kernel_ms > 0


def test_ref_index_ht32(dev_factory):
    """m/4 >= 65535 switches to 32-bit slots (lz_diff.cpp:146) and no longer fits shared memory"""
    rng = np.random.default_rng(10)
    mml = 20
    dev = dev_factory(k=25, min_match_len=mml)
    ref = rng.integers(0, 4, 300000).astype(np.uint8)
    text = mutate(rng, ref, 0.002, 4)
    dev.set_splitters(np.zeros(0, np.uint64))
    dev.scan_contigs([to_fasta_body(ref), to_fasta_body(text)])
    dev.put_references([(0, 0, len(ref), False, 20)])
    z = orc.LZ(ref, mml)
    ht, short = z.ht()
    assert not short
    assert np.array_equal(dev.get_index(20), ht)
    assert dev.lz_encode([(1, 0, len(text), False, 20)])[0] == z.encode(text)
    assert int(dev.lz_estimate([(1, 0, len(text), False, 20)])[0]) == z.estimate(text)


def test_put_reference_host(dev_factory):
    rng = np.random.default_rng(11)
    mml = 20
    dev = dev_factory(k=25, min_match_len=mml)
    for g, dirty in ((30, False), (31, True)):
        ref = rng.integers(0, 4, 5000).astype(np.uint8)
        if dirty:
            ref[100:120] = 4; ref[3000] = 11
        text = mutate(rng, ref, 0.01, 1)
        dev.set_splitters(np.zeros(0, np.uint64))
        dev.scan_contigs([to_fasta_body(text)])
        dev.put_reference_host(g, ref)
        z = orc.LZ(ref, mml)
        assert np.array_equal(dev.get_index(g), z.ht()[0])
        assert dev.lz_encode([(0, 0, len(text), False, g)])[0] == z.encode(text)


def test_preprocess_and_scan(dev_factory):
    """preprocess_raw_contig + compress_contig cut list vs the oracle, incl. non-ACGT resets and empty contigs"""
    rng = np.random.default_rng(12)
    for k in (17, 25, 31, 32):
        dev = dev_factory(k=k, min_match_len=20, segment_size=1000)
        ref = rng.integers(0, 4, 60000).astype(np.uint8)
        spl, _ = orc.determine_splitters([ref], k, 1000)
        assert len(spl) > 20
        contigs_codes = [ref, mutate(rng, ref, 0.01, 5), mutate(rng, ref, 0.001, 2, 6), np.zeros(0, np.uint8),
                         rng.integers(0, 4, k - 1).astype(np.uint8), rng.integers(0, 4, 5000).astype(np.uint8),
                         orc.revcomp(ref)]
        # plant adjacent splitter occurrences (second hit inside the k-mer refill window must be ignored)
        c = contigs_codes[1].copy(); c[3000:3000 + 200] = np.concatenate([ref[5000:5100], ref[5000:5100]]); contigs_codes[1] = c
        raws = [to_fasta_body(cc, 70, 0.2, rng) for cc in contigs_codes]
        dev.set_splitters(spl)
        cuts = dev.scan_contigs(raws)
        for ci, (cc, raw) in enumerate(zip(contigs_codes, raws)):
            pre = orc.preprocess(raw)
            assert np.array_equal(pre, cc)
            assert int(dev.contig_len[ci]) == len(pre)
            if len(pre):
                assert np.array_equal(dev.get_segment(ci, 0, len(pre)), pre)
                assert np.array_equal(dev.get_segment(ci, 0, len(pre), True), orc.revcomp(pre))
            exp = orc.scan_contig(pre, k, spl)
            got = [x for x in cuts if x.contig == ci]
            assert len(got) == len(exp), (k, ci, len(got), len(exp))
            for a, b in zip(got, exp):
                assert (a.start, a.len, a.has_front, a.has_back) == (b.start, b.len, b.has_front, b.has_back)
                if b.has_front:
                    assert (a.front_dir, a.front_rc) == (b.front_dir, b.front_rc)
                if b.has_back:
                    assert (a.back_dir, a.back_rc) == (b.back_dir, b.back_rc)


def test_preprocess_odd_bytes(dev_factory):
    """every byte value 0..127 through the cnv_num table, CRLF line ends, unaligned contig starts"""
    rng = np.random.default_rng(13)
    dev = dev_factory(k=21, min_match_len=20)
    dev.set_splitters(np.zeros(0, np.uint64))
    raws = [bytes(rng.integers(0, 128, n).astype(np.uint8)) for n in (1, 15, 16, 17, 8191, 8192, 8193, 20000)]
    raws.append(b"ACGT\r\nacgtNNNN\r\n>@`xyz\r\n")
    dev.scan_contigs(raws)
    for i, r in enumerate(raws):
        pre = orc.preprocess(r)
        assert int(dev.contig_len[i]) == len(pre)
        if len(pre):
            assert np.array_equal(dev.get_segment(i, 0, len(pre)), pre)


def test_determine_splitters(dev_factory):
    rng = np.random.default_rng(14)
    for k, seg in ((21, 500), (31, 3000)):
        dev = dev_factory(k=k, min_match_len=20, segment_size=seg)
        contigs = [rng.integers(0, 4, n).astype(np.uint8) for n in (40000, 12000, 100, k - 1, 7000)]
        contigs[0][5000:6000] = contigs[0][20000:21000]            # repeat: those k-mers are not singletons
        contigs[1] = mutate(rng, contigs[1], 0, 0, 5)               # non-ACGT resets
        exp, _ = orc.determine_splitters(contigs, k, seg)
        got = dev.determine_splitters([to_fasta_body(c, 60) for c in contigs])
        assert np.array_equal(got, exp), (k, seg, len(got), len(exp))


def test_find_new_splitters_and_rescan(dev_factory):
    """-a mode entry points: new splitters of resident contigs that share no splitter with the reference, then the scan
    of the resident batch again under the grown set (agc_compressor.cpp:2054-2082, 1187-1229)"""
    rng = np.random.default_rng(21)
    for k, seg in ((21, 700), (31, 4000)):
        dev = dev_factory(k=k, min_match_len=20, segment_size=seg, adaptive=True)
        ref = [rng.integers(0, 4, n).astype(np.uint8) for n in (30000, 9000)]
        ref[0][3000:3600] = ref[0][12000:12600]                     # duplicated reference k-mers (v_duplicated_kmers)
        spl, _ = orc.determine_splitters(ref, k, seg)
        got = dev.determine_splitters([to_fasta_body(c, 70) for c in ref])
        assert np.array_equal(got, spl)
        ref_kmers = np.sort(np.concatenate([orc.enumerate_kmers(c, k) for c in ref]))
        novel = rng.integers(0, 4, 5 * seg + 123).astype(np.uint8)
        half = novel.copy(); half[:2 * seg] = ref[0][3000:3000 + 2 * seg]      # starts with reference sequence (incl. the duplicated part)
        withn = mutate(rng, rng.integers(0, 4, 3 * seg).astype(np.uint8), 0, 0, 4)
        rep = np.tile(rng.integers(0, 4, seg // 2).astype(np.uint8), 5)         # every k-mer occurs 5 times: no singleton
        batch = [mutate(rng, ref[0], 0.01, 2), novel, half, withn, rep, rng.integers(0, 4, k - 1).astype(np.uint8), mutate(rng, ref[1], 0.002, 0)]
        cuts0 = dev.scan_contigs([to_fasta_body(c, 60) for c in batch])
        for c, codes in enumerate(batch):
            assert [(x.start, x.len) for x in cuts0 if x.contig == c] == [(x.start, x.len) for x in orc.scan_contig(codes, k, spl)]
        per = [orc.find_new_splitters(codes, k, seg, ref_kmers) for codes in batch]
        for c in (1, 2, 3, 4, 5):
            assert np.array_equal(dev.find_new_splitters([c]), np.unique(per[c])), (k, seg, c)
        assert len(per[1]) >= 4 and len(per[4]) == 0
        new = dev.find_new_splitters([1, 2, 3, 4])
        assert np.array_equal(new, np.unique(np.concatenate([per[c] for c in (1, 2, 3, 4)])))
        grown = np.union1d(spl, new)
        dev.set_splitters(grown)
        cuts1 = dev.rescan_contigs()
        for c, codes in enumerate(batch):
            exp = orc.scan_contig(codes, k, grown)
            gotc = [x for x in cuts1 if x.contig == c]
            assert [(x.start, x.len, x.has_front, x.has_back, x.front_dir, x.back_dir, x.back_rc) for x in gotc] == \
                   [(x.start, x.len, x.has_front, x.has_back, x.front_dir if x.has_front else 0, x.back_dir if x.has_back else 0, x.back_rc if x.has_back else 0) for x in exp], (k, seg, c)


def test_filtered_kmers_and_splitter_positions(dev_factory):
    """-f mode entry points: the k-mers that pass kmer_filter_t in resident ranges (orientation, symmetry, non-ACGT resets) and
    where determine_splitters / find_new_splitters found their splitters (agc_compressor.cpp:776-802, 1826-1855)"""
    rng = np.random.default_rng(22)
    for k, seg, frac in ((21, 700, 0.05), (32, 3000, 0.01), (17, 500, 0.5)):
        thr = int(float(0xFFFFFFFFFFFFFFFF) * frac)
        dev = dev_factory(k=k, min_match_len=20, segment_size=seg, adaptive=True)
        ref = [rng.integers(0, 4, n).astype(np.uint8) for n in (20000, 6000, k - 1, 0, k)]
        ref[1] = mutate(rng, ref[1], 0, 0, 6)                       # non-ACGT symbols
        pal = rng.integers(0, 4, k // 2).astype(np.uint8)
        if k % 2 == 0:                                              # a k-mer that is its own reverse complement
            ref[0][500:500 + k] = np.concatenate([pal, orc.revcomp(pal)])
        spl, singles = orc.determine_splitters(ref, k, seg)
        got = dev.determine_splitters([to_fasta_body(c, 70) for c in ref])
        assert np.array_equal(got, spl)
        exp_pos = sorted((c, p, km, last) for c, codes in enumerate(ref) for p, km, last in orc.find_splitters_pos(codes, k, seg, singles))
        assert dev.last_splitter_positions() == exp_pos, (k, seg)
        # the reference contigs are still resident: whole contigs (len clipped), inner ranges, ranges shorter than k
        ranges = [(c, 0, 0xFFFFFFFF) for c in range(len(ref))] + [(0, 137, 5000), (0, 19990, 10), (1, 3, k), (1, 100, k - 1)]
        res = dev.filtered_kmers(ranges, thr)
        for (c, s, l), fk in zip(ranges, res):
            codes = ref[c][s:s + min(l, len(ref[c]) - s)]
            assert [(f.pos, f.kmer, f.is_dir_oriented, f.is_symmetric) for f in fk] == orc.filtered_kmers(codes, k, thr), (k, c, s, l)
        # after a scan, find_new_splitters reports positions in the new batch
        novel = [rng.integers(0, 4, 4 * seg + 77).astype(np.uint8), mutate(rng, rng.integers(0, 4, 3 * seg).astype(np.uint8), 0, 0, 3)]
        dev.scan_contigs([to_fasta_body(c, 60) for c in [ref[0]] + novel])
        ref_kmers = np.sort(np.concatenate([orc.enumerate_kmers(c, k) for c in ref]))
        new = dev.find_new_splitters([1, 2])
        where = dev.last_splitter_positions()
        assert sorted(set(w[2] for w in where)) == list(new)
        for c in (1, 2):
            codes = novel[c - 1]
            km = orc.enumerate_kmers(codes, k)
            u, cnt = np.unique(km, return_counts=True)
            cand = np.setdiff1d(u[cnt == 1], ref_kmers)
            assert [w[1:] for w in where if w[0] == c] == sorted(orc.find_splitters_pos(codes, k, seg, cand)), (k, c)


def test_assign(dev_factory):
    rng = np.random.default_rng(15)
    k = 25
    dev = dev_factory(k=k, min_match_len=20, segment_size=2000)
    ref = rng.integers(0, 4, 50000).astype(np.uint8)
    spl, _ = orc.determine_splitters([ref], k, 2000)
    dev.set_splitters(spl)
    cuts = dev.scan_contigs([to_fasta_body(ref), to_fasta_body(orc.revcomp(ref))])
    a0 = dev.assign_cuts(cuts)
    keys = {}
    for c, a in zip(cuts, a0):
        if a.klass == 0:
            assert a.group_id == -1
            keys.setdefault((a.key1, a.key2), 16 + len(keys))
        if a.klass == 3:
            assert a.group_id == 0
    ks = list(keys.items())
    dev.map_insert([x[0][0] for x in ks], [x[0][1] for x in ks], [x[1] for x in ks])
    a1 = dev.assign_cuts(cuts)
    for c, a in zip(cuts, a1):
        fc = min(c.front_dir, c.front_rc); bc = min(c.back_dir, c.back_rc)
        if c.has_front and c.has_back:
            assert (a.key1, a.key2) == (min(fc, bc), max(fc, bc))
            assert a.is_rc == (0 if fc < bc else 1)
            assert a.group_id == keys[(a.key1, a.key2)]
    # forward and reverse-complement contigs must land in the same groups with opposite orientation
    fwd = sorted((a.group_id, a.is_rc) for c, a in zip(cuts, a1) if c.contig == 0 and a.klass == 0)
    rev = sorted((a.group_id, 1 - a.is_rc) for c, a in zip(cuts, a1) if c.contig == 1 and a.klass == 0)
    assert fwd == rev


def test_pack_refs(dev_factory):
    rng = np.random.default_rng(16)
    dev = dev_factory(k=25, min_match_len=20)
    refs = [rng.integers(0, 4, n).astype(np.uint8) for n in (0, 1, 2, 3, 4, 5, 17, 4096, 60031)]
    refs.append(np.tile(rng.integers(0, 4, 7).astype(np.uint8), 900))     # period 7 -> raw + level 19 (segment.h:251-254)
    refs.append(np.tile(rng.integers(0, 4, 40).astype(np.uint8), 200))    # period 40: not detected (lags 4..31 only)
    r = rng.integers(0, 4, 5001).astype(np.uint8); r[100:130] = 4; refs.append(r)                 # N  -> 3 symbols / byte
    r = rng.integers(0, 4, 5002).astype(np.uint8); r[7] = 5; r[4000:4003] = 4; refs.append(r)     # R  -> 3 / byte
    r = rng.integers(0, 4, 5003).astype(np.uint8); r[9] = 11; refs.append(r)                      # B  -> 2 / byte
    r = rng.integers(0, 4, 5004).astype(np.uint8); r[11] = 30; refs.append(r)                     # other -> raw + 0x10
    r = np.full(3000, 4, np.uint8); r[::50] = 1; refs.append(r)                                    # N-rich and periodic
    dev.set_splitters(np.zeros(0, np.uint64))
    dev.scan_contigs([to_fasta_body(r) for r in refs])
    dev.put_references([(i, 0, len(r), False, 16 + i) for i, r in enumerate(refs)])
    got = dev.pack_refs([16 + i for i in range(len(refs))], sum(len(r) + 2 for r in refs))
    for (payload, use), r in zip(got, refs):
        assert use == orc.ref_use_tuples(r), len(r)
        assert payload == (orc.bytes2tuples(r) if use else r.tobytes()), len(r)
