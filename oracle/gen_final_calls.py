"""Generate tests/golden/final_calls.npz by running the UNMODIFIED reference final-call step.

Run in the build container only:   PYTHONDONTWRITEBYTECODE=1 python oracle/gen_final_calls.py

The step after the network is prepareVcf.vcfRecords (/root/reference/python/prepareVcf.py:110-175): for every site
it calls callAlleles (:36-105) on each expert's pair probabilities, on the expert np.argmax(meta) picks ("best") and
on the float64 re-mix ("mean").  prepareVcf imports two helpers that need pysam / PyVCF (absent here); they are
replaced by stubs that only record what callAlleles hands them, so the selection rule, tie-break, QUAL formula and
re-mix that reach the golden file are the reference's own code.  Inputs are the per-expert pair probabilities and
meta weights the reference produced for the hybrid_full / single_tech golden cases (tests/golden/*.npz) plus
hand-made sites with exact ties.
"""
from __future__ import annotations

import json
import os
import pickle
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/python"
OUT = os.path.join(ROOT, "tests", "golden", "final_calls.npz")
NAMES = ["G", "GA", "T", "C", "GTT", "A"]           # allele strings of a site, first = reference allele


def install_stubs(ref_alleles):
    vcf = types.ModuleType("vcfFromContigs")

    def createVcfRecord(chromosome, start, ref, index, refAlleles, altAlleles, genotypes, string="", qual=0):
        return [json.dumps({"pos": start, "ref": refAlleles[0], "alt": altAlleles[0], "gt": genotypes[0], "qual": qual})]
    vcf.createVcfRecord = createVcfRecord
    sys.modules["vcfFromContigs"] = vcf

    fasta = types.ModuleType("PySamFastaWrapper")

    class PySamFastaWrapper:
        def __init__(self, database=None):
            self.chrom = None

        def __getitem__(self, sl):
            return list(ref_alleles[sl.start])

    fasta.PySamFastaWrapper = PySamFastaWrapper
    sys.modules["PySamFastaWrapper"] = fasta


def sites_from_golden(case):
    g = np.load(os.path.join(ROOT, "tests", "golden", case + ".npz"))
    sao = g["site_allele_off"]
    experts, meta = g["pair_experts"], g["site_meta"]
    out, p0 = [], 0
    for s in range(len(sao) - 1):
        n = int(sao[s + 1] - sao[s])
        npairs = n * (n + 1) // 2
        if n >= 2:
            out.append((n, experts[:, p0:p0 + npairs].astype(np.float32), meta[s].astype(np.float32)))
        p0 += npairs
    return out


def tie_sites():
    """Exact ties between pairs (so the key order decides) and between meta weights (np.argmax takes the first)."""
    half, quarter = np.float32(0.5), np.float32(0.25)
    e = np.array([[quarter, half, half], [half, quarter, half], [half, half, quarter]], np.float32)
    return [(2, e, np.array([0.25, 0.5, 0.25], np.float32)),
            (2, e, np.array([0.4, 0.4, 0.2], np.float32)),
            (2, np.full((3, 3), 1.0, np.float32), np.array([1.0, 0.0, 0.0], np.float32)),   # QUAL cap at p = 1
            (3, np.tile(np.array([.1, .2, .2, .2, .1, .2], np.float32), (3, 1)), np.array([0.2, 0.3, 0.5], np.float32))]


def main():
    import torch
    sites = sites_from_golden("hybrid_full") + sites_from_golden("hybrid_ensemble2") + sites_from_golden("single_tech") + tie_sites()
    ref_alleles, names_per_site, loci = {}, [], []
    for k, (n, experts, meta) in enumerate(sites):
        names = NAMES[:n] if k % 2 == 0 else NAMES[:n][::-1]       # reference allele first; vary the name order
        pos = 1000 + 10 * k
        ref_alleles[pos] = names[0]
        names_per_site.append(names)
        loci.append(("chr1", pos, len(names[0])))
    # The .features records (caller_calling.py:746-754; values are 0-d torch tensors) are made by the product's own host
    # function, so the reference's vcfRecords below consumes exactly what a hello_b200 caller would pickle.
    sys.path.insert(0, ROOT)
    from hello_b200.model import feature_records
    pair_off = np.concatenate([[0], np.cumsum([s[1].shape[1] for s in sites])])
    items = feature_records(np.concatenate([s[1] for s in sites], axis=1), np.stack([s[2] for s in sites]), pair_off,
                            names_per_site, loci)
    install_stubs(ref_alleles)
    sys.path.insert(0, REF)
    import prepareVcf                                                # the reference
    with tempfile.TemporaryDirectory() as tmp:
        data = os.path.join(tmp, "site.features")
        pickle.dump(items, open(data, "wb"))
        prepareVcf.vcfRecords(data, "unused", tmp)
        recs = {k: [json.loads(l) for l in open(data.replace(tmp, tmp) + ".%s.vcf" % k)] for k in
                ("expert0", "expert1", "expert2", "best", "mean")}
        choices = [int(l.split("\t")[3]) for l in open(data + ".choices.bed")]

    def top_pair(rec, names):
        alleles = [rec["ref"] if g == 0 else rec["alt"][g - 1] for g in rec["gt"]]
        return names.index(alleles[0]), names.index(alleles[1])

    S = len(items)
    n_alleles = np.array([s[0] for s in sites], np.int32)
    call_pair = np.zeros((S, 5, 2), np.int32)
    call_qual = np.zeros((S, 5), np.float64)
    order = np.zeros((S, 6), np.int32)                               # allele name index per allele slot (-1 pad)
    order[:] = -1
    for k, it in enumerate(items):
        n = sites[k][0]
        names = NAMES[:n] if k % 2 == 0 else NAMES[:n][::-1]
        order[k, :n] = [NAMES.index(x) for x in names]
        for c, key in enumerate(("expert0", "expert1", "expert2", "best", "mean")):
            call_pair[k, c] = top_pair(recs[key][k], names)
            call_qual[k, c] = recs[key][k]["qual"]
    P = sum(n * (n + 1) // 2 for n in n_alleles)
    experts = np.concatenate([s[1] for s in sites], axis=1)
    assert experts.shape == (3, P)
    np.savez_compressed(OUT, n_alleles=n_alleles, experts=experts, meta=np.stack([s[2] for s in sites]),
                        allele_name_idx=order, names=np.array(NAMES), call_pair=call_pair, call_qual=call_qual,
                        best_expert=np.array(choices, np.int32), calls=np.array(["expert0", "expert1", "expert2", "best", "mean"]))
    print("wrote", OUT, "sites", S, "choices", choices)


if __name__ == "__main__":
    main()
