"""Plain-Python statement of the tdna_row_result fields (include/taxondna_cuda.h), evaluated on an
ORACLE distance matrix: the expectation the CUDA row-summary kernel is compared with."""
import math

import numpy as np

NEAR = 2e-5


def row_results(D, species):
    n = D.shape[0]
    out = []
    for q in range(n):
        spq = int(species[q])
        qnamed = spq >= 0
        r = dict(best=-1, con=-1, allo=-1, other=-1, d_best=-1.0, d_con=-1.0, d_allo=-1.0, d_other=-1.0, tie_count=0,
                 tie_species_min=0x7fffffff, tie_species_max=-1, tie_unnamed=0, block_count=0, near_best=0, near_con=0,
                 near_allo=0, near_other=0, unnamed_at_con=0, unnamed_at_allo=0, unnamed_at_other=0, n_valid=0, flags=0)
        row = D[q]
        if not (row[q] < 0):
            r["flags"] |= 1
        cand = []
        for j in range(n):
            d = float(row[j])
            if j == q or d < 0:
                continue
            r["n_valid"] += 1
            if math.isnan(d) or math.isinf(d):
                r["flags"] |= 2
            if math.isnan(d):
                continue
            spj = int(species[j])
            rank = 0 if (qnamed and spj == spq) else 1
            cand.append((d, rank, j, spj))
        if cand:
            b = min(cand, key=lambda t: t[:3])
            r["best"], r["d_best"] = b[2], b[0]
            spb = b[3]
            named = [t for t in cand if t[3] >= 0]
            if qnamed:
                con = [t for t in named if t[3] == spq]
                allo = [t for t in named if t[3] != spq]
                if con:
                    c = min(con, key=lambda t: (t[0], t[2]))
                    r["con"], r["d_con"] = c[2], c[0]
                if allo:
                    a = min(allo, key=lambda t: (t[0], t[2]))
                    r["allo"], r["d_allo"] = a[2], a[0]
            others = [t for t in named if t[3] != spb and t[2] != b[2]]
            o = min(others, key=lambda t: t[:3]) if others else None
            if o:
                r["other"], r["d_other"] = o[2], o[0]
            for (d, rank, j, spj) in cand:
                if d == b[0] and j != b[2]:
                    r["tie_count"] += 1
                    if spj < 0:
                        r["tie_unnamed"] += 1
                    else:
                        r["tie_species_min"] = min(r["tie_species_min"], spj)
                        r["tie_species_max"] = max(r["tie_species_max"], spj)
                for nm in ("best", "con", "allo", "other"):
                    ref = r["d_" + nm]
                    if r[nm] >= 0:
                        if d != ref and abs(d - ref) < NEAR:
                            r["near_" + nm] += 1
                        if nm != "best" and spj < 0 and d == ref:
                            r["unnamed_at_" + nm] += 1
                if j != b[2] and spj >= 0 and spj == spb:
                    if o is None or (d, rank, j) < o[:3]:
                        r["block_count"] += 1
        out.append(r)
    return out


def compare(res, exp, shared=None):
    """res: numpy structured array from the library; exp: list of dicts."""
    for q, e in enumerate(exp):
        for k, v in e.items():
            got = res[k][q]
            if isinstance(v, float):
                assert np.float64(got).view(np.uint64) == np.float64(v).view(np.uint64), (q, k, got, v)
            else:
                assert int(got) == v, (q, k, int(got), v)
        if shared is not None:
            for nm in ("con", "allo"):
                j = e[nm]
                assert int(res["shared_" + nm][q]) == (int(shared[q, j]) if j >= 0 else 0), (q, nm)
