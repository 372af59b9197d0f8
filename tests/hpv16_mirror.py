"""Python mirror of `rkmh hpv16` (reference src/rkmh.cpp:2366-2723) on top of an ENGINE that supplies the two
hot-path pieces: raw hashes of sequences and per-read set-intersection counts.  With the oracle as engine it pins
the oracle's restatement to the real binary's golden stdout (CPU test); with the GPU context as engine it checks the
C ABI (rkg_hash_batch / rkg_load_ref_sets / rkg_classify_sets) against the same goldens."""
import numpy as np

from rkmh_b200.fastx import load_sequences


def lineage_tables(names, hashes, prefix_len):
    """{group: set of hashes} by the first `prefix_len` letters of the name, then for each group (sorted by key, as
    std::map iterates) the hashes no other group holds, ascending (src/rkmh.cpp:2560-2644)."""
    groups = {}
    for nm, h in zip(names, hashes):
        groups.setdefault(nm[:prefix_len], set()).update(int(x) for x in h)
    keys = sorted(groups)
    out = []
    for k in keys:
        own = set(groups[k])
        for o in keys:
            if o != k:
                own -= groups[o]
        out.append(np.array(sorted(own), np.uint64))
    return keys, out


class OracleEngine:
    def __init__(self, oracle):
        self.o = oracle

    def hashes(self, seqs, k):
        return [self.o.calc_hashes(self.o.to_upper(s), [k]) for s in seqs]

    def classify(self, reads, ks, sets, ntypes, counts_reads=None, min_occ=-1, table_size=800_000_000):
        counts = None
        if min_occ >= 0:
            counts = np.zeros(table_size, np.int32)
            self.o.count_batch(counts_reads, ks, counts)
        # the reference compares with ASCENDING arrays (minhashes() sorted the type arrays in place)
        return self.o.classify_sets(reads, ks, [np.sort(np.asarray(x, np.uint64)) for x in sets], ntypes, counts=counts,
                                    min_occ=max(min_occ, 0))


class GpuEngine:
    def __init__(self, gpu):
        self.g = gpu

    def hashes(self, seqs, k):
        from rkmh_b200 import pack_sequences
        self.g.set_params([k], 1)
        h, ho = self.g.calc_hashes(*pack_sequences(seqs))
        return [h[ho[i]:ho[i + 1]] for i in range(len(seqs))]

    def classify(self, reads, ks, sets, ntypes, counts_reads=None, min_occ=-1, table_size=800_000_000):
        from rkmh_b200 import pack_sequences
        self.g.set_params(ks, 1, read_table_size=table_size if min_occ >= 0 else 0)
        self.g.load_ref_sets(sets)
        if min_occ >= 0:
            self.g.count_reset()
            self.g.count_batch(*pack_sequences(counts_reads))
        return self.g.classify_sets(*pack_sequences(reads), ntypes, min_occ=min_occ)


def hpv16_stdout(engine, type_file, sub_file, read_files, ks, min_occ=-1):
    tn, ts, _ = load_sequences([type_file])
    sn, ss, _ = load_sequences([sub_file])
    rn, rs, _ = load_sequences(read_files)
    k0 = ks[0]
    type_h = engine.hashes(ts, k0)
    sub_h = engine.hashes(ss, k0)
    lin_names, lin_tabs = lineage_tables([n.decode() for n in sn], sub_h, 1)
    sub_names, sub_tabs = lineage_tables([n.decode() for n in sn], sub_h, 2)
    sets = list(type_h) + lin_tabs + sub_tabs
    hits, extra = engine.classify(rs, ks, sets, len(ts), counts_reads=rs, min_occ=min_occ)
    out = []
    for i, name in enumerate(rn):
        hashnum = sum(max(len(rs[i]) - k, 0) for k in ks)
        line = [name.decode(), tn[int(hits["ref_id"][i])].decode(), f"{int(hits['best'][i])}/{hashnum}"]
        tiers = []
        for names, cnt in ((lin_names, extra[i, :len(lin_names)]), (sub_names, extra[i, len(lin_names):])):
            order = sorted(range(len(names)), key=lambda j: -int(cnt[j]))      # stable, like std::sort on <= 16 elements
            tiers.append(("".join(f"{names[j]}:{'%g' % (int(cnt[j]) / hashnum)};" for j in order),
                          "".join(f"{int(cnt[j])};" for j in order)))
        line += [tiers[0][0], tiers[1][0], tiers[0][1], tiers[1][1]]
        out.append("\t".join(line) + "\n")
    side = "".join(n + "\t" + "".join(f"{int(x)}\t" for x in t) + "\n" for n, t in zip(lin_names, lin_tabs))
    return "".join(out).encode(), side.encode()
