"""Synthetic COI-like alignments of the BASELINE.json shapes.

Star phylogeny: root -> genus (8 % substitutions) -> species (3 %) -> individual (0.5 %),
transitions:transversions = 2:1, COI-like base composition.  Rows are generated in
(genus, species, replicate) order with fixed-width alphabetic genus/species codes and
zero-padded GI numbers, so generation order IS the reference's SORT_BYNAME order
(Sequence.compareByDisplayName, Common/DNA/Sequence.java:864-887) and full-name order.

numpy only; used by bench.py, the tests and the examples.  Not part of the compute path.
"""
import numpy as np

CONFIGS = {
    # name: (genera, species per genus, replicates, L, mode, extras)
    "C2": dict(genera=19, species=23, reps=5, L=2664, mode=0, n_frac=0.01),
    "C3": dict(genera=80, species=10, reps=10, L=658, mode=1),
    "C4": dict(genera=50, species=10, reps=10, L=2664, mode=0, gappy=True),
    "H": dict(genera=100, species=50, reps=10, L=658, mode=0),
    "C5": dict(genera=1000, species=100, reps=10, L=658, mode=0),
    # H's shape with C4's noise (8 % '?', 8 % '-' half of it external, 4 % IUPAC letters): what real barcode sets look like to the
    # count kernels (no tile is free of gaps / ambiguity letters)
    "Hg": dict(genera=100, species=50, reps=10, L=658, mode=0, gappy=True),
    "tiny": dict(genera=3, species=4, reps=5, L=658, mode=0),
}
SEED0 = 20240607
_SEED_INDEX = {"C2": 0, "C3": 1, "C4": 2, "C5": 3, "H": 4, "tiny": 5, "Hg": 6}  # fixed: adding a config must not move the others' seeds

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_TS = np.zeros(256, dtype=np.uint8)
for a, b in (b"AG", b"GA", b"CT", b"TC"):
    _TS[a] = b
_TV = np.zeros((256, 2), dtype=np.uint8)
for a, (b, c) in {"A": "CT", "G": "CT", "C": "AG", "T": "AG"}.items():
    _TV[ord(a)] = (ord(b), ord(c))
_AMBIG = np.frombuffer(b"RYKMSWBDHVN", dtype=np.uint8)


def _mutate(rng, seqs, rate):
    """seqs: uint8 [..., L]; independent per-site substitutions at `rate`, ts:tv = 2:1."""
    r = rng.random(seqs.shape, dtype=np.float32)
    hit = r < rate
    kind = rng.random(seqs.shape, dtype=np.float32)
    out = seqs.copy()
    ts = hit & (kind < 2.0 / 3.0)
    tv = hit & ~ts
    out[ts] = _TS[seqs[ts]]
    which = (kind[tv] > 5.0 / 6.0).astype(np.intp)
    out[tv] = _TV[seqs[tv], which]
    return out


def _code(i, width, lead):
    s = ""
    for _ in range(width):
        s = chr(ord("a") + i % 26) + s
        i //= 26
    return lead + s


def normalise_external_gaps(sym):
    """Vectorised Sequence.changeSequence external-gap rule (Sequence.java:806-840) for rows that
    are already upper-case IUPAC: leading/trailing runs of '-' (stepping over '?') become '_'."""
    sym = sym.copy()
    n, L = sym.shape
    solid = (sym != ord("-")) & (sym != ord("?"))
    has = solid.any(axis=1)
    first = np.where(has, solid.argmax(axis=1), L)
    last = np.where(has, L - 1 - solid[:, ::-1].argmax(axis=1), -1)
    col = np.arange(L)[None, :]
    ext = (col < first[:, None]) | (col > last[:, None])
    sym[ext & (sym == ord("-"))] = ord("_")
    return sym


def generate(config="C3", seed=None, names=True, genera=None, species=None, reps=None, L=None, scale_n=None):
    """Returns dict(symbols uint8 [N, L] (normalised alphabet), species_id, genus_id, epithet_rank,
    fullname_rank (all int32 [N]), names (list of str or None), mode, config)."""
    cfg = dict(CONFIGS[config])
    if genera is not None:
        cfg["genera"] = genera
    if species is not None:
        cfg["species"] = species
    if reps is not None:
        cfg["reps"] = reps
    if L is not None:
        cfg["L"] = L
    if seed is None:
        seed = SEED0 + _SEED_INDEX[config]
    rng = np.random.Generator(np.random.PCG64(seed))
    G, S, R, Ln = cfg["genera"], cfg["species"], cfg["reps"], cfg["L"]
    if scale_n is not None:  # weak scaling: grow the number of genera to reach ~scale_n sequences
        G = max(1, int(round(scale_n / (S * R))))
    N = G * S * R
    root = rng.choice(_ACGT, size=Ln, p=[0.30, 0.16, 0.15, 0.39])
    sym = np.empty((N, Ln), dtype=np.uint8)
    gstep = max(1, 4096 // (S * R))
    for g0 in range(0, G, gstep):
        g1 = min(G, g0 + gstep)
        gen = _mutate(rng, np.broadcast_to(root, (g1 - g0, Ln)), 0.08)
        spc = _mutate(rng, np.repeat(gen, S, axis=0), 0.03)
        ind = _mutate(rng, np.repeat(spc, R, axis=0), 0.005)
        sym[g0 * S * R:g1 * S * R] = ind
    if cfg.get("n_frac"):
        m = rng.random(sym.shape, dtype=np.float32) < cfg["n_frac"]
        sym[m] = ord("N")
    if cfg.get("gappy"):
        # 20 % of sites: 8 % '?', 8 % '-' (half of it as leading/trailing runs), 4 % IUPAC codes
        r = rng.random(sym.shape, dtype=np.float32)
        sym[r < 0.08] = ord("?")
        sym[(r >= 0.08) & (r < 0.12)] = ord("-")
        amb = (r >= 0.12) & (r < 0.16)
        sym[amb] = rng.choice(_AMBIG, size=int(amb.sum()))
        lead = rng.integers(0, int(0.04 * Ln) + 1, size=N)
        trail = rng.integers(0, int(0.04 * Ln) + 1, size=N)
        col = np.arange(Ln)[None, :]
        sym[(col < lead[:, None]) | (col >= Ln - trail[:, None])] = ord("-")
        sym = normalise_external_gaps(sym)
    gi = np.arange(N)
    genus_id = (gi // (S * R)).astype(np.int32)
    species_id = (gi // R).astype(np.int32)
    out = dict(symbols=sym, species_id=species_id, genus_id=genus_id, epithet_rank=species_id.copy(),
               fullname_rank=gi.astype(np.int32), mode=cfg["mode"], config=config, names=None,
               shape=(G, S, R, Ln))
    if names:
        wg = max(4, len(str(G)))
        out["names"] = ["gi|%09d| %s %s" % (i + 1, _code(i // (S * R), wg, "G"), _code(i // R, wg + 1, "s")) for i in range(N)]
    return out
