#!/usr/bin/env python
"""Regenerates tests/golden/* from the reference's own data and documentation.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

Inputs (read-only): the three genomes under /root/reference/testdata/old/ and the
known answers printed in /root/reference/README.md and
/root/reference/analysis/distance/README.md (K1..K9 of SURVEY.md section 4).
Outputs:
  * genomes.npz   -- MG1655 and IAI39 packed 2 bits/base (A0 C1 G2 T3; both genomes are
                     ACGT-only single records) so the -m gpu tests can run K1/K2/K4..K7/K9
                     on the GPU box, where /root/reference does not exist;
  * kat.json      -- the documented known answers plus oracle-derived digests
                     (count / xor / sum mod 2^64 of each sorted result) that were
                     checked against those known answers when this script ran.
"""
import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

REF = "/root/reference/testdata/old"
GENOMES = {
    "mg1655": "Ecoli-MG1655.fasta.gz",
    "iai39": "Ecoli-IAI39.fasta.gz",
    "amuc": "A.muciniphila-ATCC_BAA-835.fasta.gz",
}

# Known answers, with the reference line that prints each.
KAT = {
    "K1_mg1655_k23_canonical_unique": 4546632,   # README.md:156,203-204
    "K2_iai39_k23_canonical_unique": 4902266,    # README.md:201-202
    "K3_amuc_k23_canonical_unique": 2630905,     # README.md:171,200
    "K4_union": 6872728,                         # README.md:215,277-278
    "K5_inter": 2576170,                         # README.md:240,276
    "K6_diff_iai39_minus_mg1655": 2326096,       # README.md:246,270
    "K7_first3_sorted_mg1655": ["AAAAAAAAACCATCCAAATCTGG", "AAAAAAAAACCGCTAGTATATTC",
                                "AAAAAAAAACCTGAAAAAAACGG"],  # README.md:177-180
    "K8_nthash_k23_canonical": {"CATCCGCCATCTTTGGGGTGTCG": 1210726578792,
                                "AGCGCAAAATCCCCAAACATGTA": 2286899379883,
                                "AACTGATTTTTGATGATGACTCC": 3542156397282},  # README.md:183-186
    "K9_mg1655_k31_nthash_scaled15": 586734,     # analysis/distance/README.md:9,16,44
    "K12_mg1655_k31_minimizer_w15": 549963,      # analysis/distance/README.md:8,15,39 (count -k 31 -K -H -W 15)
}


def read_fasta_records(path):
    """bio/seqio/fastx view of a FASTA file: list of sequences, line breaks stripped."""
    recs, cur = [], []
    with gzip.open(path, "rb") as fh:
        for line in fh:
            line = line.rstrip(b"\r\n")
            if line.startswith(b">"):
                if cur:
                    recs.append(b"".join(cur))
                cur = []
            elif line:
                cur.append(line)
    if cur:
        recs.append(b"".join(cur))
    return recs


def pack2(seq: bytes) -> np.ndarray:
    lut = np.full(256, 255, dtype=np.uint8)
    for i, c in enumerate(b"ACGT"):
        lut[c] = i
    v = lut[np.frombuffer(seq, dtype=np.uint8)]
    assert (v < 4).all(), "genome is not ACGT-only"
    pad = (-len(v)) % 4
    v = np.concatenate([v, np.zeros(pad, dtype=np.uint8)]).reshape(-1, 4)
    return (v[:, 0] | (v[:, 1] << 2) | (v[:, 2] << 4) | (v[:, 3] << 6)).astype(np.uint8)


def unpack2(packed: np.ndarray, n: int) -> bytes:
    v = np.empty((len(packed), 4), dtype=np.uint8)
    for i in range(4):
        v[:, i] = (packed >> (2 * i)) & 3
    return np.frombuffer(b"ACGT", dtype=np.uint8)[v.reshape(-1)[:n]].tobytes()


def digest(a: np.ndarray) -> dict:
    a = np.asarray(a, dtype=np.uint64)
    return {"n": int(len(a)), "xor": int(np.bitwise_xor.reduce(a)) if len(a) else 0,
            "sum": int(np.add.reduce(a, dtype=np.uint64)) if len(a) else 0}


def main():
    seqs = {}
    for name, fn in GENOMES.items():
        recs = read_fasta_records(os.path.join(REF, fn))
        assert len(recs) == 1, (name, len(recs))
        seqs[name] = recs[0]

    sets = {}
    for name, s in seqs.items():
        off = np.array([0, len(s)], dtype=np.uint64)
        sets[name] = oracle.count(s, off, 23, canonical=True, hashed=False)
    assert len(sets["mg1655"]) == KAT["K1_mg1655_k23_canonical_unique"]
    assert len(sets["iai39"]) == KAT["K2_iai39_k23_canonical_unique"]
    assert len(sets["amuc"]) == KAT["K3_amuc_k23_canonical_unique"]
    a, b = sets["iai39"], sets["mg1655"]  # glob order of the README run: IAI39 first
    u, _ = oracle.union([a, b])
    i, _ = oracle.inter([a, b])
    d, _ = oracle.diff([a, b])
    assert len(u) == KAT["K4_union"] and len(i) == KAT["K5_inter"] and len(d) == KAT["K6_diff_iai39_minus_mg1655"]
    assert [oracle.decode(int(c), 23).decode() for c in sets["mg1655"][:3]] == KAT["K7_first3_sorted_mg1655"]
    for kmer, h in KAT["K8_nthash_k23_canonical"].items():
        assert int(oracle.nthash_iter(kmer.encode(), 23, canonical=True)[0]) == h
    mg = seqs["mg1655"]
    off = np.array([0, len(mg)], dtype=np.uint64)
    max_hash = int(float(2**64 - 1) / 15.0)  # count.go:98: uint64(float64(^uint64(0)) / float64(scale))
    hs = oracle.count(mg, off, 31, canonical=True, hashed=True)
    sc = oracle.count(mg, off, 31, canonical=True, hashed=True, scaled=True, max_hash=max_hash)
    assert len(sc) == KAT["K9_mg1655_k31_nthash_scaled15"], len(sc)
    mz = oracle.count_minimizer(mg, off, 31, 15, canonical=True)
    assert len(mz) == KAT["K12_mg1655_k31_minimizer_w15"], len(mz)

    out = dict(KAT)
    out["max_hash_scale15"] = max_hash
    out["digests"] = {
        "mg1655_k23": digest(sets["mg1655"]), "iai39_k23": digest(sets["iai39"]),
        "union": digest(u), "inter": digest(i), "diff": digest(d),
        "mg1655_k31_nthash": digest(hs), "mg1655_k31_nthash_scaled15": digest(sc), "mg1655_k31_minimizer_w15": digest(mz),
        "mg1655_k31_kmer_noncanonical": digest(oracle.count(mg, off, 31, canonical=False, hashed=False)),
        "mg1655_k21_circular": digest(oracle.count(mg, off, 21, canonical=True, hashed=False, circular=True)),
    }
    out["lengths"] = {k: len(v) for k, v in seqs.items()}
    with open(os.path.join(HERE, "kat.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "genomes.npz"),
                        mg1655=pack2(seqs["mg1655"]), mg1655_len=len(seqs["mg1655"]),
                        iai39=pack2(seqs["iai39"]), iai39_len=len(seqs["iai39"]))
    # round-trip check of the packing
    z = np.load(os.path.join(HERE, "genomes.npz"))
    assert unpack2(z["mg1655"], int(z["mg1655_len"])) == seqs["mg1655"]
    print("golden fixtures written:", json.dumps(out["digests"], indent=1))


if __name__ == "__main__":
    main()
