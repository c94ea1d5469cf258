"""Pure-Python restatement of the reference's distance *consumers* -- TEST INFRASTRUCTURE ONLY.

Paths cited below are relative to
/root/reference/src/main/java/com/ggvaidya/TaxonDNA/.

Parity status: the reference stores no expected output for any of these
consumers, so this file is pinned only by reading the Java
source: "parity unpinned" by the reference's own tests.  What it *can* be checked
against is the one hand-derivable case (the Tests/best match fixture: both
queries "No match.") -- see tests/test_oracle_consumers.py.

Everything here works on list positions (ints) over an `OList`, with distances
supplied by a `Dist` object that wraps full numpy matrices from the C oracle
(or, in parity tests, matrices produced by the CUDA library -- that is the only
way product output ever enters this file, and only inside tests/).
"""
import functools
import math
import re

import numpy as np

from . import PDM_K2P, PDM_TRANS_ONLY, PDM_UNCORRECTED, all_pairs, lib, normalise, pack_rows  # noqa: F401


# --------------------------------------------------------------------------
# Java number formatting
# --------------------------------------------------------------------------
def _java_sci(digits, exp10):
    """digits: significant decimal digits string (no dot), value = 0.d1d2.. * 10^exp10"""
    # Java: d.ddd E n  where n = exp10 - 1
    mant = digits[0] + "." + (digits[1:] if len(digits) > 1 else "0")
    return mant + "E" + str(exp10 - 1)


def _shortest_digits(x, single=False):
    """(digits, exp10) with value = 0.DIGITS * 10**exp10, shortest round-trip."""
    if single:
        s = np.format_float_scientific(np.float32(x), unique=True, trim="-")
    else:
        s = np.format_float_scientific(np.float64(x), unique=True, trim="-")
    mant, e = s.split("e")
    mant = mant.replace("-", "")
    digits = mant.replace(".", "")
    exp10 = int(e) + 1
    digits = digits.rstrip("0") or "0"
    return digits, exp10


def java_double_str(x, single=False):
    """Double.toString / Float.toString (JDK >= 19 shortest-digits behaviour)."""
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Infinity" if x > 0 else "-Infinity"
    sign = "-" if math.copysign(1.0, x) < 0 else ""
    ax = abs(x)
    if ax == 0:
        return sign + "0.0"
    digits, exp10 = _shortest_digits(ax, single)
    if 1e-3 <= ax < 1e7:
        if exp10 <= 0:
            body = "0." + "0" * (-exp10) + digits
        elif exp10 >= len(digits):
            body = digits + "0" * (exp10 - len(digits)) + ".0"
        else:
            body = digits[:exp10] + "." + digits[exp10:]
        return sign + body
    return sign + _java_sci(digits, exp10)


def java_float_str(x):
    return java_double_str(np.float32(x), single=True)


# Settings.* (Common/DNA/Settings.java:69-97) -- via the C oracle
def make_long(d):
    return lib.tdo_make_long(float(d))


def percentage(x, y):
    return lib.tdo_percentage(float(x), float(y))


def identical(d1, d2):
    return bool(lib.tdo_settings_identical(float(d1), float(d2)))


# --------------------------------------------------------------------------
# Sequence (names only; symbols are normalised by the C oracle)
# --------------------------------------------------------------------------
_P3 = re.compile(r"([A-Z][a-z]+) ([a-z]+) ([a-z]+)\b")
_P2 = re.compile(r"([A-Z][a-z]+) ([a-z]+)\b")
_PSP = re.compile(r"([A-Z][a-z]+) ([a-z]+)\.\b")
_PGI = re.compile(r"gi\|(\d+)[\|:]")
_PFAM = re.compile(r"\(family:\s*([A-Za-z]+)\s*\)")


class OSeq:
    """Restates Sequence's naming logic (Common/DNA/Sequence.java:668-742, 228-271, 446-471)."""

    def __init__(self, name, raw):
        self.change_name(name)
        self.seq = normalise(raw)  # raises ValueError like SequenceException
        self.length = len(self.seq)

    def change_name(self, name):
        name = name.replace("\n", " ").strip()
        self.name = name
        self.genus = self.species = self.subspecies = self.gi = self.family = ""
        self.warning = False
        m = _P3.search(name)
        if m:
            self.genus, self.species, self.subspecies = m.group(1), m.group(2), m.group(3)
        else:
            m = _P2.search(name)
            if m:
                self.genus, self.species = m.group(1), m.group(2)
            else:
                m = _PSP.search(name)
                if m:
                    self.genus, self.species = m.group(1), m.group(2)
                    self.warning = True
        if self.genus == "":
            self.warning = True
        m = _PGI.search(name)
        if m:
            self.gi = m.group(1)
        m = _PFAM.search(name)
        if m:
            self.family = m.group(1)

    def species_name(self):
        if not self.species or not self.genus:
            return None
        return self.genus + " " + self.species

    def get_gi(self):
        return self.gi or None

    def display_name(self):
        if self.warning:
            return "{" + self.name[:80] + "}"
        if self.get_gi() is None:
            return self.name
        sp = self.species_name()
        if sp is not None:
            n = sp
            if self.subspecies:
                n += " (" + self.subspecies + ")"
            if self.family:
                n += " (Family " + self.family + ")"
            n += " (gi:" + self.gi + ")"
            return n
        return self.name


def read_fasta(path):
    """FastaFile.appendFromFile + makeSequence (Common/DNA/formats/FastaFile.java:144-289).

    Sequences the reference would demote to BaseSequence (invalid symbols) raise.
    `path` may end in .gz (the committed golden fixtures are gzipped)."""
    import gzip
    if str(path).endswith(".gz"):
        with gzip.open(path, "rt", encoding="latin-1") as fh:
            return read_fasta_text(fh.read())
    with open(path, "r", encoding="latin-1") as fh:
        return read_fasta_text(fh.read())


def read_fasta_text(text):
    out = []
    name, seq = "", ""

    def flush():
        if name:
            nm = name.replace("_", " ")
            s = seq.replace("U", "T").replace("u", "t") if ("U" in seq or "u" in seq) else seq
            out.append(OSeq(nm, re.sub(r"\s+", "", s)))

    for line in text.splitlines():
        if re.match(r"^\s*$", line):
            continue
        if re.match(r"^\s*#.*$", line):
            continue
        m = re.match(r"^>\s*(.+)\s*$", line)
        if m:
            flush()
            name, seq = m.group(1), ""
            continue
        m = re.match(r"^>\s*(.*)\s*$", line)
        if m:
            flush()
            name, seq = "No name specified in file", ""
            continue
        m = re.match(r"^\s*(.*)\s*$", line)
        seq += m.group(1)
    flush()
    return out


def _cmp_ignore_case(a, b):
    """String.compareToIgnoreCase: per char upper-then-lower, then length."""
    for ca, cb in zip(a, b):
        if ca != cb:
            ua, ub = ca.upper(), cb.upper()
            if ua != ub:
                la, lb = ua.lower(), ub.lower()
                if la != lb:
                    return ord(la[0]) - ord(lb[0])
    return len(a) - len(b)


def compare_by_display_name(s1, s2):
    """Sequence.compareByDisplayName (Common/DNA/Sequence.java:864-887)."""
    if s2.warning and not s1.warning:
        return -1
    if not s2.warning and s1.warning:
        return 1
    return _cmp_ignore_case(s1.display_name(), s2.display_name())


def sort_by_name(seqs):
    """SequenceList.resort(SORT_BYNAME): stable sort (Common/DNA/SequenceList.java:198-205, 337-380)."""
    return sorted(seqs, key=functools.cmp_to_key(compare_by_display_name))


# --------------------------------------------------------------------------
# Distances over a list
# --------------------------------------------------------------------------
class Config:
    def __init__(self, mode=PDM_UNCORRECTED, min_overlap=300, ambig=True):
        self.mode, self.min_overlap, self.ambig = mode, min_overlap, ambig


class Dist:
    """Sequence.getPairwise / getSharedLength over list positions.

    `dist`/`shared` may be injected (parity tests pass the CUDA library's matrices)."""

    def __init__(self, seqs, cfg, dist=None, shared=None, threads=1):
        self.seqs, self.cfg = seqs, cfg
        if dist is None:
            rows, lens = pack_rows([s.seq for s in seqs])
            r = all_pairs(rows, lens, cfg.mode, cfg.min_overlap, cfg.ambig, threads=threads)
            dist, shared = r["dist"], r["shared"]
        self.dist, self.shared = dist, shared

    def pairwise(self, i, j):
        return float(self.dist[i, j])

    def shared_length(self, i, j):
        return int(self.shared[i, j])


# --------------------------------------------------------------------------
# java.util.TimSort, restated from the published algorithm (Peters' listsort as
# ported to the JDK: MIN_MERGE 32, MIN_GALLOP 7, binary insertion for short runs,
# galloping merges, and the post-JDK-8072909 mergeCollapse invariant).
# Collections.sort(sorted, cmp) at SortedSequenceList.java:120 lands here via
# List.sort -> Arrays.sort(Object[], Comparator).
# --------------------------------------------------------------------------
class ComparisonContractError(Exception):
    """IllegalArgumentException("Comparison method violates its general contract!")"""


class _TimSort:
    MIN_MERGE = 32
    MIN_GALLOP = 7

    def __init__(self, a, c):
        self.a, self.c = a, c
        self.min_gallop = self.MIN_GALLOP
        self.run_base, self.run_len = [], []

    @classmethod
    def sort(cls, a, c):
        lo, hi = 0, len(a)
        n = hi - lo
        if n < 2:
            return
        if n < cls.MIN_MERGE:
            init = cls.count_run(a, lo, hi, c)
            cls.binary_sort(a, lo, hi, lo + init, c)
            return
        ts = cls(a, c)
        min_run = cls.min_run_length(n)
        while True:
            run_len = cls.count_run(a, lo, hi, c)
            if run_len < min_run:
                force = n if n <= min_run else min_run
                cls.binary_sort(a, lo, lo + force, lo + run_len, c)
                run_len = force
            ts.run_base.append(lo)
            ts.run_len.append(run_len)
            ts.merge_collapse()
            lo += run_len
            n -= run_len
            if n == 0:
                break
        ts.merge_force_collapse()

    @staticmethod
    def binary_sort(a, lo, hi, start, c):
        if start == lo:
            start += 1
        while start < hi:
            pivot = a[start]
            left, right = lo, start
            while left < right:
                mid = (left + right) >> 1
                if c(pivot, a[mid]) < 0:
                    right = mid
                else:
                    left = mid + 1
            a[left + 1:start + 1] = a[left:start]
            a[left] = pivot
            start += 1

    @staticmethod
    def count_run(a, lo, hi, c):
        run_hi = lo + 1
        if run_hi == hi:
            return 1
        if c(a[run_hi], a[lo]) < 0:
            run_hi += 1
            while run_hi < hi and c(a[run_hi], a[run_hi - 1]) < 0:
                run_hi += 1
            a[lo:run_hi] = a[lo:run_hi][::-1]
        else:
            run_hi += 1
            while run_hi < hi and c(a[run_hi], a[run_hi - 1]) >= 0:
                run_hi += 1
        return run_hi - lo

    @classmethod
    def min_run_length(cls, n):
        r = 0
        while n >= cls.MIN_MERGE:
            r |= n & 1
            n >>= 1
        return n + r

    def merge_collapse(self):
        rl = self.run_len
        while len(rl) > 1:
            n = len(rl) - 2
            if (n > 0 and rl[n - 1] <= rl[n] + rl[n + 1]) or (n > 1 and rl[n - 2] <= rl[n] + rl[n - 1]):
                if rl[n - 1] < rl[n + 1]:
                    n -= 1
            elif n < 0 or rl[n] > rl[n + 1]:
                break
            self.merge_at(n)

    def merge_force_collapse(self):
        rl = self.run_len
        while len(rl) > 1:
            n = len(rl) - 2
            if n > 0 and rl[n - 1] < rl[n + 1]:
                n -= 1
            self.merge_at(n)

    def merge_at(self, i):
        a, c = self.a, self.c
        base1, len1 = self.run_base[i], self.run_len[i]
        base2, len2 = self.run_base[i + 1], self.run_len[i + 1]
        self.run_len[i] = len1 + len2
        del self.run_base[i + 1]
        del self.run_len[i + 1]
        k = self.gallop_right(a[base2], a, base1, len1, 0, c)
        base1 += k
        len1 -= k
        if len1 == 0:
            return
        len2 = self.gallop_left(a[base1 + len1 - 1], a, base2, len2, len2 - 1, c)
        if len2 == 0:
            return
        if len1 <= len2:
            self.merge_lo(base1, len1, base2, len2)
        else:
            self.merge_hi(base1, len1, base2, len2)

    @staticmethod
    def gallop_left(key, a, base, length, hint, c):
        last_ofs, ofs = 0, 1
        if c(key, a[base + hint]) > 0:
            max_ofs = length - hint
            while ofs < max_ofs and c(key, a[base + hint + ofs]) > 0:
                last_ofs = ofs
                ofs = (ofs << 1) + 1
            if ofs > max_ofs:
                ofs = max_ofs
            last_ofs += hint
            ofs += hint
        else:
            max_ofs = hint + 1
            while ofs < max_ofs and c(key, a[base + hint - ofs]) <= 0:
                last_ofs = ofs
                ofs = (ofs << 1) + 1
            if ofs > max_ofs:
                ofs = max_ofs
            last_ofs, ofs = hint - ofs, hint - last_ofs
        last_ofs += 1
        while last_ofs < ofs:
            m = last_ofs + ((ofs - last_ofs) >> 1)
            if c(key, a[base + m]) > 0:
                last_ofs = m + 1
            else:
                ofs = m
        return ofs

    @staticmethod
    def gallop_right(key, a, base, length, hint, c):
        ofs, last_ofs = 1, 0
        if c(key, a[base + hint]) < 0:
            max_ofs = hint + 1
            while ofs < max_ofs and c(key, a[base + hint - ofs]) < 0:
                last_ofs = ofs
                ofs = (ofs << 1) + 1
            if ofs > max_ofs:
                ofs = max_ofs
            last_ofs, ofs = hint - ofs, hint - last_ofs
        else:
            max_ofs = length - hint
            while ofs < max_ofs and c(key, a[base + hint + ofs]) >= 0:
                last_ofs = ofs
                ofs = (ofs << 1) + 1
            if ofs > max_ofs:
                ofs = max_ofs
            last_ofs += hint
            ofs += hint
        last_ofs += 1
        while last_ofs < ofs:
            m = last_ofs + ((ofs - last_ofs) >> 1)
            if c(key, a[base + m]) < 0:
                ofs = m
            else:
                last_ofs = m + 1
        return ofs

    def merge_lo(self, base1, len1, base2, len2):
        a, c = self.a, self.c
        tmp = a[base1:base1 + len1]
        cursor1, cursor2, dest = 0, base2, base1
        a[dest] = a[cursor2]
        dest += 1
        cursor2 += 1
        len2 -= 1
        if len2 == 0:
            a[dest:dest + len1] = tmp[cursor1:cursor1 + len1]
            return
        if len1 == 1:
            a[dest:dest + len2] = a[cursor2:cursor2 + len2]
            a[dest + len2] = tmp[cursor1]
            return
        min_gallop = self.min_gallop
        done = False
        while not done:
            count1 = count2 = 0
            while True:
                if c(a[cursor2], tmp[cursor1]) < 0:
                    a[dest] = a[cursor2]
                    dest += 1
                    cursor2 += 1
                    count2 += 1
                    count1 = 0
                    len2 -= 1
                    if len2 == 0:
                        done = True
                        break
                else:
                    a[dest] = tmp[cursor1]
                    dest += 1
                    cursor1 += 1
                    count1 += 1
                    count2 = 0
                    len1 -= 1
                    if len1 == 1:
                        done = True
                        break
                if (count1 | count2) >= min_gallop:
                    break
            if done:
                break
            while True:
                count1 = self.gallop_right(a[cursor2], tmp, cursor1, len1, 0, c)
                if count1 != 0:
                    a[dest:dest + count1] = tmp[cursor1:cursor1 + count1]
                    dest += count1
                    cursor1 += count1
                    len1 -= count1
                    if len1 <= 1:
                        done = True
                        break
                a[dest] = a[cursor2]
                dest += 1
                cursor2 += 1
                len2 -= 1
                if len2 == 0:
                    done = True
                    break
                count2 = self.gallop_left(tmp[cursor1], a, cursor2, len2, 0, c)
                if count2 != 0:
                    a[dest:dest + count2] = a[cursor2:cursor2 + count2]
                    dest += count2
                    cursor2 += count2
                    len2 -= count2
                    if len2 == 0:
                        done = True
                        break
                a[dest] = tmp[cursor1]
                dest += 1
                cursor1 += 1
                len1 -= 1
                if len1 == 1:
                    done = True
                    break
                min_gallop -= 1
                if not (count1 >= self.MIN_GALLOP or count2 >= self.MIN_GALLOP):
                    break
            if done:
                break
            if min_gallop < 0:
                min_gallop = 0
            min_gallop += 2
        self.min_gallop = 1 if min_gallop < 1 else min_gallop
        if len1 == 1:
            a[dest:dest + len2] = a[cursor2:cursor2 + len2]
            a[dest + len2] = tmp[cursor1]
        elif len1 == 0:
            raise ComparisonContractError()
        else:
            a[dest:dest + len1] = tmp[cursor1:cursor1 + len1]

    def merge_hi(self, base1, len1, base2, len2):
        a, c = self.a, self.c
        tmp = a[base2:base2 + len2]
        cursor1 = base1 + len1 - 1
        cursor2 = len2 - 1
        dest = base2 + len2 - 1
        a[dest] = a[cursor1]
        dest -= 1
        cursor1 -= 1
        len1 -= 1
        if len1 == 0:
            a[dest - (len2 - 1):dest + 1] = tmp[0:len2]
            return
        if len2 == 1:
            dest -= len1
            cursor1 -= len1
            a[dest + 1:dest + 1 + len1] = a[cursor1 + 1:cursor1 + 1 + len1]
            a[dest] = tmp[cursor2]
            return
        min_gallop = self.min_gallop
        done = False
        while not done:
            count1 = count2 = 0
            while True:
                if c(tmp[cursor2], a[cursor1]) < 0:
                    a[dest] = a[cursor1]
                    dest -= 1
                    cursor1 -= 1
                    count1 += 1
                    count2 = 0
                    len1 -= 1
                    if len1 == 0:
                        done = True
                        break
                else:
                    a[dest] = tmp[cursor2]
                    dest -= 1
                    cursor2 -= 1
                    count2 += 1
                    count1 = 0
                    len2 -= 1
                    if len2 == 1:
                        done = True
                        break
                if (count1 | count2) >= min_gallop:
                    break
            if done:
                break
            while True:
                count1 = len1 - self.gallop_right(tmp[cursor2], a, base1, len1, len1 - 1, c)
                if count1 != 0:
                    dest -= count1
                    cursor1 -= count1
                    len1 -= count1
                    a[dest + 1:dest + 1 + count1] = a[cursor1 + 1:cursor1 + 1 + count1]
                    if len1 == 0:
                        done = True
                        break
                a[dest] = tmp[cursor2]
                dest -= 1
                cursor2 -= 1
                len2 -= 1
                if len2 == 1:
                    done = True
                    break
                count2 = len2 - self.gallop_left(a[cursor1], tmp, 0, len2, len2 - 1, c)
                if count2 != 0:
                    dest -= count2
                    cursor2 -= count2
                    len2 -= count2
                    a[dest + 1:dest + 1 + count2] = tmp[cursor2 + 1:cursor2 + 1 + count2]
                    if len2 <= 1:
                        done = True
                        break
                a[dest] = a[cursor1]
                dest -= 1
                cursor1 -= 1
                len1 -= 1
                if len1 == 0:
                    done = True
                    break
                min_gallop -= 1
                if not (count1 >= self.MIN_GALLOP or count2 >= self.MIN_GALLOP):
                    break
            if done:
                break
            if min_gallop < 0:
                min_gallop = 0
            min_gallop += 2
        self.min_gallop = 1 if min_gallop < 1 else min_gallop
        if len2 == 1:
            dest -= len1
            cursor1 -= len1
            a[dest + 1:dest + 1 + len1] = a[cursor1 + 1:cursor1 + 1 + len1]
            a[dest] = tmp[cursor2]
        elif len2 == 0:
            raise ComparisonContractError()
        else:
            a[dest - (len2 - 1):dest + 1] = tmp[0:len2]


def java_sort(a, cmp):
    """In-place Arrays.sort(Object[], Comparator) (TimSort)."""
    _TimSort.sort(a, cmp)


def _java_long_to_int(v):
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


# --------------------------------------------------------------------------
# SortedSequenceList (Common/DNA/SortedSequenceList.java:78-128, 197-261)
# --------------------------------------------------------------------------
def sort_against(D, q, order=None):
    """Returns the list positions that survive `sortAgainst(query)`, in sorted order."""
    seqs = D.seqs
    n = len(seqs)
    cand = [j for j in (order if order is not None else range(n)) if not (D.pairwise(j, q) < 0)]
    qname = seqs[q].species_name()

    def compare(o1, o2):
        d1, d2 = D.pairwise(q, o1), D.pairwise(q, o2)
        if d1 < 0:
            return 1
        if d2 < 0:
            return -1
        diff = make_long(d1 - d2)
        if diff == 0:
            if o1 == q and o2 == q:
                return 0
            elif o1 == q:
                return -1
            elif o2 == q:
                return 1
            n1, n2 = seqs[o1].species_name(), seqs[o2].species_name()
            if n1 is None or n2 is None:
                return 0
            if n1 == qname and n2 == qname:
                return 0
            if n1 == qname:
                return -1
            elif n2 == qname:
                return 1
            return 0
        return _java_long_to_int(diff)

    java_sort(cand, compare)
    return cand


class JavaNPE(Exception):
    """Stands for the NullPointerException the reference would die with."""


# --------------------------------------------------------------------------
# BestMatch.run (SpeciesIdentifier/BestMatch.java:164-582)
# --------------------------------------------------------------------------
def best_match(D, threshold_percent=3.0):
    """Returns (text, per_query records).  threshold_percent is the text-field value."""
    seqs = D.seqs
    n = len(seqs)
    threshold = float(threshold_percent)
    threshold /= 100
    cnt = dict(no_names=0, bm_correct=0, bm_ambiguous=0, bm_incorrect=0, bcm_correct=0, bcm_ambiguous=0,
               bcm_incorrect=0, bcm_nomatch=0, allo_at_zero=0, zero_pct=0, valid_consp=0, no_matches=0)
    listing = ["Query\tMatch\tIdentification\n"]
    records = []
    for q in range(n):
        query = seqs[q]
        if query.species_name() is None:
            cnt["no_names"] += 1
            records.append(None)
            continue
        ss = sort_against(D, q)
        count_sequences = len(ss)
        best = ss[1] if count_sequences > 1 else None
        listing.append(query.display_name())
        if best is None or D.pairwise(best, q) == -1:
            listing.append("\t\tNo match.\n")
            cnt["no_matches"] += 1
            records.append(dict(best=-1))
            continue
        bd = D.pairwise(best, q)
        if identical(bd, 0):
            cnt["zero_pct"] += 1
        clean_block = mixed_block = False
        count_best = 0
        for y in range(2, count_sequences):
            match = ss[y]
            if not identical(D.pairwise(match, q), bd):
                break
            count_best += 1
            if seqs[match].species_name() is None:
                raise JavaNPE("BestMatch.java:319 match.getSpeciesName() is null")
            if seqs[match].species_name() == seqs[best].species_name():
                clean_block = True
            else:
                mixed_block = True
        first_con = first_allo = None
        for y in range(1, count_sequences):
            match = ss[y]
            mname = seqs[match].species_name()
            if mname is None:
                continue
            if D.pairwise(match, q) >= 0:
                if first_con is None and mname == query.species_name():
                    first_con = match
                    cnt["valid_consp"] += 1
                elif first_allo is None and mname != query.species_name():
                    first_allo = match
                if first_con is not None and first_allo is not None:
                    break
        if first_con is None:
            listing.append("\tNo conspecific in database\t---\t0")
        else:
            listing.append("\t" + seqs[first_con].display_name() + "\t" +
                           java_double_str(percentage(D.pairwise(q, first_con), 1)) + "\t" +
                           str(D.shared_length(q, first_con)))
        if first_allo is None:
            listing.append("\tNo allospecific in database\t---\t0")
        else:
            listing.append("\t" + seqs[first_allo].display_name() + "\t" +
                           java_double_str(percentage(D.pairwise(q, first_allo), 1)) + "\t" +
                           str(D.shared_length(q, first_allo)))
        conspecific = seqs[best].species_name() is not None and seqs[best].species_name() == query.species_name()
        pct = java_double_str(percentage(bd, 1))
        within = bd <= threshold
        tail = " (within threshold)\n" if within else " (outside threshold)\n"
        if not clean_block and not mixed_block:
            listing.append("\t" + seqs[best].display_name())
            verdict = None
            if conspecific:
                listing.append("\tSuccessful match at " + pct + "%")
                cnt["bm_correct"] += 1
                cnt["bcm_correct" if within else "bcm_nomatch"] += 1
            else:
                if identical(bd, 0):
                    cnt["allo_at_zero"] += 1
                listing.append("\tIncorrect match at " + pct + "%")
                cnt["bm_incorrect"] += 1
                cnt["bcm_incorrect" if within else "bcm_nomatch"] += 1
            listing.append(tail)
        elif clean_block and not mixed_block:
            listing.append("\t" + seqs[best].display_name() + " and " + str(count_best) + " others")
            if conspecific:
                listing.append("\tSuccessful match at " + pct + "%")
                cnt["bm_correct"] += 1
                cnt["bcm_correct" if within else "bcm_nomatch"] += 1
            else:
                if identical(bd, 0):
                    cnt["allo_at_zero"] += 1
                listing.append("\tIncorrect match at " + pct + "%")
                cnt["bm_incorrect"] += 1
                cnt["bcm_incorrect" if within else "bcm_nomatch"] += 1
            listing.append(tail)
        else:
            if identical(bd, 0):
                cnt["allo_at_zero"] += 1
            listing.append("\t" + seqs[best].display_name() + " and " + str(count_best) +
                           " others from different species\tMultiple species found at " + pct +
                           "%, identification with certainty is impossible")
            cnt["bm_ambiguous"] += 1
            cnt["bcm_ambiguous" if within else "bcm_nomatch"] += 1
            listing.append(tail)
        records.append(dict(best=best, d_best=bd, count_best=count_best, clean=clean_block, mixed=mixed_block,
                            first_con=first_con, first_allo=first_allo))
    valid = n - cnt["no_matches"] - cnt["no_names"]

    def pc(x, y):
        return java_double_str(percentage(x, y))

    text = ("Sequences:\t" + str(n) + "\n"
            "Sequences without recognizable species names (ignored in all subsequent counts):\t" + str(cnt["no_names"]) +
            "\nSequences with atleast one matching sequence in the data set:\t" + str(valid) + "\n"
            "Sequences with atleast one matching conspecific sequence in the data set:\t" + str(cnt["valid_consp"]) +
            "\nSequences with a closest match at 0%:\t" + str(cnt["zero_pct"]) +
            "\nAllospecific matches at 0%:\t" + str(cnt["allo_at_zero"]) + "\t(" + pc(cnt["allo_at_zero"], cnt["zero_pct"]) +
            "% of all matches at 0%)"
            "\n\nCorrect identifications according to \"Best Match\":\t" + str(cnt["bm_correct"]) + " (" + pc(cnt["bm_correct"], valid) + "%)"
            "\nAmbiguous according to \"Best Match\":\t" + str(cnt["bm_ambiguous"]) + " (" + pc(cnt["bm_ambiguous"], valid) + "%)"
            "\nIncorrect identifications according to \"Best Match\":\t" + str(cnt["bm_incorrect"]) + " (" + pc(cnt["bm_incorrect"], valid) + "%)"
            "\n\nCorrect identifications according to \"Best Close Match\":\t" + str(cnt["bcm_correct"]) + " (" + pc(cnt["bcm_correct"], valid) + "%)"
            "\nAmbiguous according to \"Best Close Match\":\t" + str(cnt["bcm_ambiguous"]) + " (" + pc(cnt["bcm_ambiguous"], valid) + "%)"
            "\nIncorrect identifications according to \"Best Close Match\":\t" + str(cnt["bcm_incorrect"]) + " (" + pc(cnt["bcm_incorrect"], valid) + "%)"
            "\nSequences without any match closer than " + pc(threshold, 1) + "%:\t" + str(cnt["bcm_nomatch"]) + " (" + pc(cnt["bcm_nomatch"], valid) + "%)"
            "\n\n" + "".join(listing))
    return text, records, cnt


# --------------------------------------------------------------------------
# SpeciesDetail(s) (Common/DNA/SpeciesDetail.java:78-95; SpeciesDetails.java:60-120)
# --------------------------------------------------------------------------
def species_valid_conspecific_counts(D):
    """name -> getSequencesWithValidConspecificsCount()."""
    seqs, cfg = D.seqs, D.cfg
    groups = {}
    for i, s in enumerate(seqs):
        nm = s.species_name()
        if nm is not None:
            groups.setdefault(nm, []).append(i)
    out = {}
    for nm, members in groups.items():
        valid = 0
        for a in members:
            for b in members:
                if a == b:
                    continue
                if D.shared_length(a, b) >= cfg.min_overlap:
                    valid += 1
                    break
        out[nm] = valid
    return out, {nm: len(m) for nm, m in groups.items()}


# --------------------------------------------------------------------------
# BlockAnalysis.run (SpeciesIdentifier/BlockAnalysis.java:179-387)
# --------------------------------------------------------------------------
def block_analysis(D, threshold_percent=3.0):
    seqs = D.seqs
    n = len(seqs)
    threshold = float(threshold_percent) / 100
    c = dict(no_name=0, correct=0, incorrect=0, ambiguous=0, nomatch=0, noseq=0)
    listing = ["Query\tFound in a complete block?\n"]
    nonames = ["Sequences without an identifiable species name\n"]
    valid_consp, _ = species_valid_conspecific_counts(D)
    records = []
    for q in range(n):
        query = seqs[q]
        name_query = query.species_name()
        if name_query is None:
            c["no_name"] += 1
            nonames.append(query.name + "\n")
            records.append(None)
            continue
        ss = sort_against(D, q)
        block_size = 1
        is_correct = is_incorrect = nomatches = False
        name_first = ""
        if len(ss) > 0:
            if len(ss) < 2:
                raise JavaNPE("BlockAnalysis.java:275 sset.get(1) is null")
            name_first = seqs[ss[1]].species_name()
        if len(ss) == 0:
            c["noseq"] += 1
        elif D.pairwise(ss[1], q) > threshold:
            nomatches = True
        else:
            for y in range(2, len(ss)):
                mname = seqs[ss[y]].species_name()
                if mname is None:
                    continue
                if mname != name_first:
                    break
                block_size += 1
            if block_size <= 1:
                c["ambiguous"] += 1
            else:
                if name_first is None:
                    raise JavaNPE("BlockAnalysis.java:299 getSpeciesDetailsByName(null)")
                vc = valid_consp[name_first]
                if block_size == vc:
                    if name_first != name_query:
                        c["incorrect"] += 1
                        is_incorrect = True
                    else:
                        c["ambiguous"] += 1
                elif block_size == vc - 1:
                    if name_first == name_query:
                        c["correct"] += 1
                        is_correct = True
                    else:
                        c["ambiguous"] += 1
                else:
                    c["ambiguous"] += 1
        listing.append(query.name)
        if is_correct:
            listing.append("\tcorrect\t")
        elif is_incorrect:
            listing.append("\tincorrect\t")
        elif nomatches:
            listing.append("\tno match\t")
            c["nomatch"] += 1
        else:
            listing.append("\tambiguous\t")
        listing.append("\n")
        records.append(dict(block_size=block_size, correct=is_correct, incorrect=is_incorrect, nomatch=nomatches))
    valid = n - c["noseq"]

    def pc(x, y):
        return java_double_str(percentage(x, y))

    text = ("Sequences:\t" + str(n) + "\nSequences without valid species names:\t" + str(c["no_name"]) +
            "\nSequences with atleast one sequence with an overlap of " + str(D.cfg.min_overlap) + " base pairs:\t" + str(valid) +
            "\n\nCorrect identifications according to \"All Species Barcodes\":\t" + str(c["correct"]) + " (" + pc(c["correct"], valid) + "%)"
            "\nAmbiguous according to \"All Species Barcodes\":\t" + str(c["ambiguous"]) + " (" + pc(c["ambiguous"], valid) + "%)"
            "\nIncorrect identifications according to \"All Species Barcodes\":\t" + str(c["incorrect"]) + " (" + pc(c["incorrect"], valid) + "%)"
            "\nSequences with no match closer than " + pc(threshold, 1) + "%:\t" + str(c["nomatch"]) + " (" + pc(c["nomatch"], valid) + "%)"
            "\n\n" + "".join(listing) + "\n\n" + "".join(nonames))
    return text, records, c


# --------------------------------------------------------------------------
# PairwiseDistribution (Common/DNA/PairwiseDistribution.java:77-203, 247-297, 357-396)
# --------------------------------------------------------------------------
PD_INTRA, PD_INTER = 0, 1


class PairwiseDistribution:
    """`seqs` must already be in SORT_BYNAME order (conspecificIterator re-sorts, SequenceList.java:616)."""

    def __init__(self, D, kind):
        seqs = D.seqs
        n = len(seqs)
        vals = []
        self.count_sequences = n
        if kind == PD_INTRA:
            first_idx = {}
            for i, s in enumerate(seqs):  # conspecificIterator: first occurrence, contiguous run
                nm = s.species_name()
                if nm is not None and nm not in first_idx:
                    first_idx[nm] = i
            for q in range(n):
                nm = seqs[q].species_name()
                if nm is None:
                    continue
                x = first_idx[nm]
                while x < n and seqs[x].species_name() is not None and seqs[x].species_name() == nm:
                    if x != q and not (seqs[x].name < seqs[q].name):
                        d = np.float32(D.pairwise(q, x))
                        if d != np.float32(-1):
                            vals.append(d)
                    x += 1
        else:
            for q in range(n):
                a = seqs[q]
                for j in range(n):
                    if j == q:
                        continue
                    b = seqs[j]
                    if a.genus == b.genus and a.species != b.species and a.species < b.species:
                        d = np.float32(D.pairwise(q, j))
                        if d != np.float32(-1):
                            vals.append(d)
        self.distances = np.sort(np.array(vals, dtype=np.float32))  # Arrays.sort(float[]): NaN last
        self.size = len(vals)

    def count_valid(self):
        return self.size

    def minimum(self):
        return np.float32(self.distances[0]) if self.size > 0 else np.float32(0)

    def maximum(self):
        return np.float32(self.distances[-1]) if self.size > 0 else np.float32(0)

    def distances_between(self, d_from, d_to):
        out = []
        count = False
        d_from, d_to = float(d_from), float(d_to)
        for v in self.distances:
            fv = float(v)
            if fv >= d_from:
                count = True
            if fv > d_to:
                count = False
            if count:
                out.append(np.float32(v))
        return out

    def between_incl(self, f_from, f_to):
        return len(self.distances_between(float(np.float32(f_from)), float(np.float32(f_to))))

    def distribution_as_string(self, d_from, d_to, d_interval, forward):
        rows = []

        def entry(label, freq, perc, cumul):
            rows.append("%s\t%s\t%s\t%s\t\n" % (label, freq, perc, cumul))

        entry("Distances", "Freq.", "Perc.", "Cumulative")
        count = self.size
        f32 = np.float32

        def ratio(f):  # (float) f_from / count in Java float arithmetic; 0/0 -> NaN
            with np.errstate(all="ignore"):
                return f32(f) / f32(count)

        f_from = self.between_incl(0, d_from)
        cumul = ratio(f_from)
        if not forward:
            cumul = f32(1.0) - cumul
        entry("<= " + java_double_str(percentage(d_from, 1.0)) + "%", f_from,
              java_double_str(percentage(f_from, count)), java_double_str(percentage(float(cumul), 1.0)))
        x = d_from + 0.000001
        while x < d_to:
            d_next = x + d_interval
            f_from = self.between_incl(x, d_next)
            cumul = f32(cumul + ratio(f_from)) if forward else f32(cumul - ratio(f_from))
            entry(java_double_str(percentage(x, 1)) + "% to " + java_double_str(percentage(d_next, 1)) + "%", f_from,
                  java_double_str(percentage(f_from, count)), java_double_str(percentage(float(cumul), 1.0)))
            x += d_interval
        f_from = self.between_incl(d_to + 0.000001, 1.0)
        cumul = f32(cumul + ratio(f_from)) if forward else f32(cumul - ratio(f_from))
        entry("> " + java_double_str(percentage(d_to, 1.0)) + "%", f_from,
              java_double_str(percentage(f_from, count)), java_double_str(percentage(float(cumul), 1.0)))
        return "".join(rows)


def _pad(x, y=None):
    buff = " " * (30 - len(x)) if len(x) < 30 else ""
    if y is None:
        return x + buff + "\n"
    return x + buff + "\t" + str(y) + "\n"


def pairwise_distances(D, kind):
    """PairwiseDistances (Common/DNA/PairwiseDistances.java:49-185) + PairwiseDistance.compareTo
    (Common/DNA/PairwiseDistance.java:58-63).  `D.seqs` in SORT_BYNAME order (the constructor re-sorts for
    PD_INTRA, :82-83).  Returns (triples (name A, name B, d) in Collections.sort order, {full name: mean})."""
    seqs = D.seqs
    n = len(seqs)
    pairs, averages = [], {}
    first_idx = {}
    for i, s in enumerate(seqs):  # conspecificIterator: first occurrence, contiguous run (SequenceList.java:614-658)
        nm = s.species_name()
        if nm is not None and nm not in first_idx:
            first_idx[nm] = i
    for q in range(n):
        a = seqs[q]
        total, count = 0.0, 0
        if kind == PD_INTRA:
            nm = a.species_name()
            if nm is None:  # :133: no pushes, no average entry
                continue
            x = first_idx[nm]
            run = []
            while x < n and seqs[x].species_name() is not None and seqs[x].species_name() == nm:
                run.append(x)
                x += 1
            others = [x for x in run if x != q]
        else:  # :160-181: same genus, different epithet, BOTH directions (the half-table test is commented out)
            others = [j for j in range(n) if j != q and a.genus == seqs[j].genus and a.species != seqs[j].species]
        for j in others:
            d = D.pairwise(q, j)
            if not (d < 0):  # distances_push (:64-68): NaN is kept
                pairs.append((q, j, d))
            if d > -1:
                total += d
                count += 1
        averages[a.name] = (total / count) if count else float("nan")  # 0.0 / 0 in Java

    def cmp(p1, p2):  # (int) Settings.makeLongFromDouble(d1 - d2)
        return _java_long_to_int(make_long(p1[2] - p2[2]))

    if pairs:
        java_sort(pairs, cmp)
    return [(seqs[q].name, seqs[j].name, d) for q, j, d in pairs], averages


def pairwise_summary(D):
    """PairwiseSummary.run (SpeciesIdentifier/PairwiseSummary.java:109-485).

    Returns (text, fivePercentCutoff as np.float32, intra, inter)."""
    seqs = D.seqs
    f32 = np.float32
    intra = PairwiseDistribution(D, PD_INTRA)
    inter = PairwiseDistribution(D, PD_INTER)
    species = {s.species_name() for s in seqs if s.species_name() is not None}
    out = []
    five = f32(0)
    mo = str(D.cfg.min_overlap)

    def pc(x, y=1):
        return java_double_str(percentage(float(x), float(y)))

    out.append(_pad("COUNTS"))
    out.append(_pad("No of sequences:", str(len(seqs)) + "\tsequences"))
    out.append(_pad("No of species:", str(len(species)) + "\tspecies"))
    out.append(_pad("Note that minimum overlap is set to " + mo + " bp: pairs of sequences with less that " + mo +
                    " bp overlap will not be counted as a pairwise distance for this analysis."))
    out.append(_pad("\nRANGES"))
    out.append(_pad("WARNING: These calculations are new to Species Identifier as of June 14, 2012, and have not been"
                    " comprehensively tested yet. Please use with caution!\n"))

    def trim_intra(dl):
        mx = dl[-1]
        x = 1
        while x < len(dl):
            dist = dl[len(dl) - x]
            if float(f32(x) / f32(len(dl))) > 0.05:
                break
            mx = dist
            x += 1
        return f32(mx)

    def trim_inter(dl, start):
        mn = start
        x = 0
        for dist in dl:
            if float(f32(x) / f32(len(dl))) > 0.05:
                break
            mn = dist
            x += 1
        return f32(mn)

    if intra.count_valid() == 0:
        out.append(_pad("There are no valid intraspecific pairwise distances (i.e. no two sequences have less than " + mo +
                        " bp overlap)."))
    else:
        dl = intra.distances_between(0, 1)
        if not dl:
            raise JavaNPE("PairwiseSummary.java:204 get(-1)")
        mx = trim_intra(dl)
        out.append(_pad("Intraspecific pairwise distances range between " + pc(intra.minimum()) + "% and " +
                        pc(intra.maximum()) + "% (contains " +
                        pc(intra.between_incl(intra.minimum(), intra.maximum()), intra.count_valid()) +
                        "% of all valid intraspecific pairwise distances)"))
        out.append(_pad("  Removing the largest 5% of intraspecific distances reduces that range to " +
                        pc(intra.minimum()) + "% and " + pc(mx) + "% (contains " +
                        pc(intra.between_incl(intra.minimum(), mx), intra.count_valid()) +
                        "% of all valid intraspecific pairwise distances)"))
    if inter.count_valid() == 0:
        out.append(_pad("There are no valid interspecific pairwise distances (i.e. no two sequences have less than " + mo +
                        " bp overlap)."))
    else:
        dl = inter.distances_between(0, 1)
        mn = trim_inter(dl, inter.minimum())
        out.append(_pad("Interspecific, intrageneric pairwise distances range between " + pc(inter.minimum()) +
                        "% and " + pc(inter.maximum()) + "% (contains " +
                        pc(inter.between_incl(inter.minimum(), inter.maximum()), inter.count_valid()) +
                        "% of all valid interspecific, intrageneric pairwise distances)"))
        out.append(_pad("  Removing the largest 5% of interspecific, intrageneric distances reduces that range to " +
                        pc(inter.minimum()) + "% and " + pc(mn) + "% (contains " +
                        pc(inter.between_incl(mn, inter.maximum()), inter.count_valid()) +
                        "% of all valid interspecific, intrageneric pairwise distances)"))
    out.append(_pad("\nPERCENTAGES"))
    min_inter = inter.minimum()
    max_intra = intra.maximum()
    count_comparisons = inter.count_valid() + intra.count_valid()
    overlap = f32(abs(f32(max_intra - min_inter)))
    if inter.count_valid() == 0:
        out.append(_pad("Total overlap:\tNo interspecific, intrageneric comparisons.\n"
                        "              \tIntraspecific comparisons range from " + pc(intra.minimum()) + "% to " +
                        pc(intra.maximum()) + "% (total width: " + pc(f32(intra.maximum() - intra.minimum())) + "%)"))
    else:
        within = intra.between_incl(min_inter, max_intra) + inter.between_incl(min_inter, max_intra)
        dl_intra = intra.distances_between(0, 1)
        dl_inter = inter.distances_between(0, 1)
        if not dl_intra:
            out.append(_pad("Total overlap:\t No intraspecific distances present."))
            out.append(_pad("Overlap with 5% error margs on both ends:\t No intraspecific distances present."))
        elif not dl_inter:
            out.append(_pad("Total overlap:\t No interspecific distances present."))
            out.append(_pad("Overlap with 5% error margs on both ends:\t No interspecific distances present."))
        else:
            out.append(_pad("Total overlap:\t" + pc(overlap) + "% (from " + pc(min_inter) + "% to " + pc(max_intra) +
                            "%, covering " + pc(within, count_comparisons) +
                            "% of all intra and interspecific but intrageneric sequences)"))
            min_inter = trim_inter(dl_inter, min_inter)
            x = 1
            while x < len(dl_intra):
                dist = dl_intra[len(dl_intra) - x]
                if float(f32(x) / f32(len(dl_intra))) > 0.05:
                    break
                max_intra = f32(dist)
                x += 1
            overlap = f32(abs(f32(max_intra - min_inter)))
            within = inter.between_incl(min_inter, max_intra) + intra.between_incl(min_inter, max_intra)
            five = f32(percentage(float(max_intra), 1))
            out.append(_pad("Overlap with 5% error margins on both ends:\t " + pc(overlap) + "% (from " + pc(min_inter) +
                            "% to " + pc(max_intra) + "%, covering " + pc(within, count_comparisons) +
                            "% of intra and interspecific but intrageneric sequences)"))
            out.append(_pad("Five percent intraspecific cutoff:\t " + java_float_str(five) + "%"))
    out.append(_pad("\nDISTRIBUTION FOR ALL INTRASPECIFIC DISTANCES (FROM 0.0 TO 0.20)"))
    out.append(_pad(intra.distribution_as_string(0.0, 0.20, 0.005, False)))
    out.append(_pad("\nDISTRIBUTION FOR ALL INTRASPECIFIC DISTANCES (FROM 0.0 TO 0.30) AT 1% intervals"))
    out.append(_pad(intra.distribution_as_string(0.0, 0.30, 0.01, False)))
    out.append(_pad("\nDISTRIBUTION FOR ALL INTRASPECIFIC DISTANCES"))
    out.append(_pad(intra.distribution_as_string(0.0, 1.0, 0.05, False)))
    out.append(_pad("\nDISTRIBUTION FOR ALL INTERSPECIFIC DISTANCES WITHIN A GENUS (FROM 0.0 to 0.20)"))
    out.append(_pad(inter.distribution_as_string(0.0, 0.20, 0.005, True)))
    out.append(_pad("\nDISTRIBUTION FOR ALL INTERSPECIFIC DISTANCES WITHIN A GENUS (FROM 0.0 TO 0.30) AT 1% intervals"))
    out.append(_pad(inter.distribution_as_string(0.0, 0.30, 0.01, True)))
    out.append(_pad("\nDISTRIBUTION FOR ALL INTERSPECIFIC DISTANCES WITHIN A GENUS"))
    out.append(_pad(inter.distribution_as_string(0.0, 1.0, 0.05, True)))
    return "".join(out), five, intra, inter


# --------------------------------------------------------------------------
# Cluster.run + the numeric part of writeupItemStrings
# (SpeciesIdentifier/Cluster.java:741-867, 188-260)
# --------------------------------------------------------------------------
def cluster(D, max_pairwise):
    """Returns (clusters as lists of list positions in the reference's order, per-cluster stats)."""
    n = len(D.seqs)
    clusters = []
    for s in range(n):
        acc = None
        ci = len(clusters)
        while ci > 0:
            ci -= 1
            cur = clusters[ci]
            k = len(cur)
            while k > 0:
                k -= 1
                cmpi = cur[k]
                d = D.pairwise(s, cmpi)
                if d < 0:
                    continue
                if d <= max_pairwise:
                    if acc is None:
                        cur.append(s)
                        acc = cur
                        k = 0
                    else:
                        acc.extend(cur)
                        del clusters[ci]
                        k = 0
        if acc is None:
            clusters.append([s])
    stats = []
    for b in clusters:
        largest = 0.0
        valid = over = 0
        for a in b:
            for c in b:
                if D.seqs[a].name == D.seqs[c].name:
                    continue
                valid += 1
                d = D.pairwise(a, c)
                if d < 0:
                    continue
                if d > largest:
                    largest = d
                if d > max_pairwise:
                    over += 1
        stats.append(dict(size=len(b), largest=largest, valid=valid, over=over,
                          pct=(percentage(over, valid) if valid != 0 else 0.0)))
    return clusters, stats


# --------------------------------------------------------------------------
# ExtremePairwise.run (SpeciesIdentifier/ExtremePairwise.java:72-210)
# --------------------------------------------------------------------------
def extreme_pairwise(D):
    """Per query: the LAST valid conspecific and the FIRST valid congeneric, allospecific entry of sortAgainst(query).
    Returns (text, records); a record is (largest-intra list position or None, smallest-inter list position or None)."""
    seqs = D.seqs
    out = ["Sequence name\tLargest conspecific match\tDistance\tOverlap\tClosest congeneric, interspecific match\tDistance\tOverlap\n"]
    records = []
    for q, seq in enumerate(seqs):
        if seq.species_name() is None:  # :109-112
            out.append(seq.name + "\tUnable to identify species name\n")
            records.append(None)
            continue
        ss = sort_against(D, q)
        largest_intra = smallest_inter = None
        for x in ss:  # :121-146
            s2 = seqs[x]
            if s2.species_name() is None:
                continue
            if seq.genus and s2.genus and seq.genus == s2.genus and seq.species_name() != s2.species_name():
                if D.pairwise(x, q) != -1:
                    smallest_inter = x
                    break
        for k in range(len(ss) - 1, 0, -1):  # :149-163: never entry 0
            x = ss[k]
            s2 = seqs[x]
            if s2.species_name() is None:
                continue
            if seq.species_name() == s2.species_name():
                if D.pairwise(x, q) != -1:
                    largest_intra = x
                    break
        out.append(seq.display_name() + "\t")
        if largest_intra is None:
            out.append("No matching conspecific sequence\tN/A\tN/A\t")
        else:
            out.append(seqs[largest_intra].display_name() + "\t" + java_double_str(percentage(D.pairwise(largest_intra, q), 1)) + "\t" +
                       str(D.shared_length(largest_intra, q)) + "\t")
        if smallest_inter is None:
            out.append("No matching congeneric, interspecific sequence\tN/A\tN/A\t")
        else:
            out.append(seqs[smallest_inter].display_name() + "\t" + java_double_str(percentage(D.pairwise(smallest_inter, q), 1)) + "\t" +
                       str(D.shared_length(smallest_inter, q)) + "\t")
        out.append("\n")
        records.append((largest_intra, smallest_inter))
    return "".join(out), records


# --------------------------------------------------------------------------
# QuerySequence.query (SpeciesIdentifier/QuerySequence.java:113-180)
# --------------------------------------------------------------------------
def query_sequence(seqs, cfg, raw_query, threads=1):
    """A pasted sequence against the list: the entries of list_result, in order.  Returns (lines, list positions)."""
    q = OSeq("Query", "".join(ch for ch in raw_query if not ch.isspace()))  # :127-133
    rows, lens = pack_rows([s.seq for s in seqs] + [q.seq])
    r = all_pairs(rows, lens, cfg.mode, cfg.min_overlap, cfg.ambig, threads=threads)
    n = len(seqs)
    D = Dist(seqs + [q], cfg, dist=r["dist"], shared=r["shared"])
    ss = sort_against(D, n, order=range(n))  # the query is not in the list: no entry is "the query" for the comparator
    lines, positions = [], []
    for x in ss:
        d = D.pairwise(n, x)
        if d < 0:
            continue
        lines.append("(%.4f%%) %s" % (d * 100, seqs[x].display_name()))  # MessageFormat "({0,number,##0.0000}%) {1}"
        positions.append(x)
    return lines, positions


# --------------------------------------------------------------------------
# PairwiseExplorer.run (SpeciesIdentifier/PairwiseExplorer.java:109-190)
# --------------------------------------------------------------------------
def pairwise_explorer(D, kind):
    """The distance categories of list_distances: [(key, [(name A, name B, d), ...])] in 0.5 % steps, then "Averages"."""
    triples, _ = pairwise_distances(D, kind)
    cats = []
    for x in range(-5, 1000, 5):
        lo, hi = x / 1000 + 0.000001, (x + 5) / 1000
        members, count = [], False  # PairwiseDistances.getDistancesBetween (:266-285): one scan of the sorted vector
        for t in triples:
            if t[2] >= lo:
                count = True
            if t[2] > hi:
                break
            if count:
                members.append(t)
        if members:
            key = java_float_str(np.float32(x) / np.float32(10)) + "% to " + java_float_str((np.float32(x) + np.float32(5)) / np.float32(10)) + \
                "% (" + str(len(members)) + " matches)"
            if x + 5 == 0:
                key = "Before 0% (" + str(len(members)) + " matches)"
            cats.append((key, members))
    cats.append(("Averages", []))
    return cats


# --------------------------------------------------------------------------
# DistanceAnalysis.run (SpeciesIdentifier/DistanceAnalysis.java:161-320)
# --------------------------------------------------------------------------
def distance_analysis(D):
    """Per genus: the mean over species of (mean over sequences of the smallest / of the mean congeneric, allospecific
    distance).  The reference iterates Hashtables keyed by Sequence objects (identity hash: the order differs from run to
    run) and by species name; sums here are taken in list order / order of first appearance, which only moves the last
    ulp of a sum that is then rounded to two decimals of a percentage (Settings.percentage)."""
    seqs = D.seqs
    if any(s.species_name() is None for s in seqs):
        raise JavaNPE("DistanceAnalysis.java:196 Hashtable.put(null, ...) / :205 getSpeciesName().equals on null")

    def average(v):
        tot, cnt = 0.0, 0
        for val in v:
            if val > -1:
                tot += val
                cnt += 1
        return tot / cnt if cnt else float("nan")

    species_order, per_seq = [], []
    for q, query in enumerate(seqs):
        if query.species_name() not in species_order:
            species_order.append(query.species_name())
        closest, inters = -1.0, []
        for x in sort_against(D, q):
            cmpseq = seqs[x]
            if cmpseq.species_name() == query.species_name():
                continue
            if cmpseq.genus == query.genus:
                dist = D.pairwise(x, q)
                if closest < 0:
                    closest = dist
                inters.append(dist)
        if closest != -1:
            per_seq.append((q, average(inters), closest))
    genera_all, genera_smallest = {}, {}
    for sp in species_order:
        genus = None
        d_all, d_small = [], []
        for q, avg, closest in per_seq:
            if seqs[q].species_name() == sp:
                genus = seqs[q].genus
                d_all.append(avg)
                d_small.append(closest)
        if genus is None:
            genus = sp[:sp.index(" ")]
        if d_all:
            genera_all.setdefault(genus, []).append(average(d_all))
            genera_smallest.setdefault(genus, []).append(average(d_small))
        else:
            genera_all.setdefault(genus, []).append(-1.0)
            genera_smallest.setdefault(genus, []).append(-1.0)
    out = ["Genus\t\t\tAvg. Avg. Smallest Inter\t\tAvg. Avg. Avg. Inter\n"]
    rows = {}
    for genus in sorted(genera_all, key=lambda g: g.encode("utf-16-be")):  # Collections.sort of Strings
        avg_inter, avg_smallest = average(genera_all[genus]), average(genera_smallest[genus])
        rows[genus] = (avg_smallest, avg_inter)
        if avg_inter < 0:
            out.append(genus + "\t\t\tNo comparisions\t\tNo comparisions\n")
        else:
            out.append(genus + "\t\t\t" + java_double_str(percentage(avg_smallest, 1)) + "\t\t" + java_double_str(percentage(avg_inter, 1)) + "\n")
    return "".join(out), rows


# --------------------------------------------------------------------------
# Cluster.makeConsensusOfBin (SpeciesIdentifier/Cluster.java:123-138): the left fold of Sequence.getConsensus
# (Common/DNA/Sequence.java:1374-1398) with consensus() (:1008-1050), getint (:1125-1143) and getcode (:1149-1196)
# --------------------------------------------------------------------------
def _getint(ch):
    v = 0
    if ch in "ARMNDHVW":
        v |= 1
    if ch in "CYMSNBHV":
        v |= 2
    if ch in "TYKNBDHW":
        v |= 4
    if ch in "GRKSNBDV":
        v |= 8
    return v


def _getcode(v):
    a, c, t, g = bool(v & 1), bool(v & 2), bool(v & 4), bool(v & 8)
    if a:
        if c:
            if t:
                return "N" if g else "H"
            return "V" if g else "M"
        if t:
            return "D" if g else "W"
        return "R" if g else "A"
    if c:
        if t:
            return "B" if g else "Y"
        return "S" if g else "C"
    if t:
        return "K" if g else "T"
    return "G" if g else "-"


def _consensus_char(ch1, ch2):
    if ch1 == "?" or ch2 == "?":
        return "?"
    if ch1 == "_":
        return "_" if ch2 == "_" else ("-" if ch2 == "-" else ch2)
    if ch2 == "_":
        return "-" if ch1 == "-" else ch1
    if ch1 in "-_":
        return ch2
    if ch2 in "-_":
        return ch1
    return _getcode(_getint(ch1) | _getint(ch2))


def get_consensus(seq1, seq2):
    """Sequence.getConsensus on two normalised sequences (bytes) -> the normalised consensus (the Sequence constructor re-normalises)."""
    a, b = seq1.decode("latin-1"), seq2.decode("latin-1")
    out = []
    for x in range(max(len(a), len(b))):
        out.append(_consensus_char(a[x] if x < len(a) else "?", b[x] if x < len(b) else "?"))
    return normalise("".join(out).replace("_", "-"))


def consensus_of_bin(seqs):
    cons = None
    for s in seqs:
        cons = s if cons is None else get_consensus(cons, s)
    return cons



# --------------------------------------------------------------------------
# SequenceMatrix's two distance callers: many small per-gene alignments
# --------------------------------------------------------------------------
class OGrid:
    """The (taxon, charset) -> Sequence grid the callers read through TableManager (SequenceMatrix/TableManager.java:254-371):
    `charsets` and `taxa` in display order, cells[(charset, taxon)] = OSeq (absent = no sequence)."""

    def __init__(self, charsets, taxa, cells):
        self.charsets, self.taxa, self.cells = list(charsets), list(taxa), dict(cells)

    def sequence_list_by_column(self, col):  # TableManager.getSequenceListByColumn :359-371
        return [(t, self.cells[(col, t)]) for t in self.taxa if (col, t) in self.cells]


def find_distances(grid, cfg, d_from, d_to, col_to_use=None, threads=1):
    """FindDistances.displayResults (SequenceMatrix/FindDistances.java:123-289): every sequence of the chosen columns against every
    other one, pairs with from <= getPairwise <= to kept in BOTH directions, Collections.sort by PairwiseDistance.compareTo.
    Returns (text of text_main, [(outer row, inner row, d)] in output order); a row is an index into the (column, taxon) enumeration."""
    cols = list(grid.charsets) if col_to_use is None else [col_to_use]  # :168-176
    rows = [(c, t, s) for c in cols for (t, s) in grid.sequence_list_by_column(c)]
    results = []
    if rows:
        D = Dist([s for (_, _, s) in rows], cfg, threads=threads)
        for a in range(len(rows)):  # :200-244 (outer column, outer sequence, inner column, inner sequence)
            for b in range(len(rows)):
                if a == b:  # Sequence.equals compares the UUIDs (Sequence.java:495-503)
                    continue
                d = D.pairwise(a, b)
                if d_from <= d <= d_to:  # :227 (NaN fails)
                    results.append((a, b, d))

    def cmp(p1, p2):  # PairwiseDistance.compareTo (Common/DNA/PairwiseDistance.java:58-63)
        return _java_long_to_int(make_long(p1[2] - p2[2]))

    java_sort(results, cmp)  # :249
    buff = []
    for (a, b, d) in results:  # :251-262: the copies were renamed "<display name> (from <column>)" (:231-235)
        na = (rows[a][2].display_name() + " (from " + rows[a][0] + ")").replace("\n", " ").strip()
        nb = (rows[b][2].display_name() + " (from " + rows[b][0] + ")").replace("\n", " ").strip()
        buff.append(na + "\t" + nb + "\t" + java_double_str(percentage(d, 1)) + "%\n")
    method = {PDM_UNCORRECTED: "uncorrected pairwise distances", PDM_K2P: "K2P distances", PDM_TRANS_ONLY: "transversions only"}.get(cfg.mode, "(unknown)")
    text = (str(len(results)) + " sequence pairs found with distances between " + java_double_str(percentage(d_from, 1)) + "% and " +
            java_double_str(percentage(d_to, 1)) + "%.\nNote that distances were calculated using " + method +
            " and sequence overlaps of less than " + str(cfg.min_overlap) + " bp weren't counted.\n\n" + "".join(buff))
    return text, results


DIST_ILLEGAL, DIST_NO_COMPARE_SEQ, DIST_SEQ_NA, DIST_NO_OVERLAP, DIST_SEQ_ON_TOP, DIST_CANCELLED = -32.0, -64.0, -128.0, -256.0, -1024.0, -2048.0


def display_distances(grid, mode, ref_taxon, selected_col, cancelled=(), threads=1):
    """DisplayDistancesMode.getSortedSequences (SequenceMatrix/DisplayDistancesMode.java:217-398): per column, every taxon's
    sequence against the reference taxon's (minOverlap forced to 1, :163), normalised per column, taxa sorted by their mean
    normalised distance over the other columns, per-column ranks.
    Returns dict(columns, taxa, distances, norm_distances, scores, ranks) with [column][taxon] tables in the sorted taxon order."""
    cfg = Config(mode, 1, True)
    columns = sorted(grid.charsets)  # getSortedColumns :196-213
    if selected_col in columns:
        columns.remove(selected_col)
        columns.insert(0, selected_col)
    taxa = list(grid.taxa)
    # one Dist over all cells: a row per (column, taxon)
    index, seqs = {}, []
    for c in columns:
        for t in taxa:
            if (c, t) in grid.cells:
                index[(c, t)] = len(seqs)
                seqs.append(grid.cells[(c, t)])
    D = Dist(seqs, cfg, threads=threads) if seqs else None
    nx, ny = len(columns), len(taxa)
    dist = [[0.0] * ny for _ in range(nx)]
    mx, mn = [-1.0] * nx, [2.0] * nx
    for x, c in enumerate(columns):  # pass 1 :255-290
        for y, t in enumerate(taxa):
            if (c, ref_taxon) not in index:
                d = DIST_NO_COMPARE_SEQ
            elif t == ref_taxon:
                d = DIST_SEQ_ON_TOP
            elif (c, t) not in index:
                d = DIST_CANCELLED if (c, t) in cancelled else DIST_SEQ_NA
            else:
                d = D.pairwise(index[(c, t)], index[(c, ref_taxon)])
                if d < 0:
                    d = DIST_NO_OVERLAP
            dist[x][y] = d
            if d >= 0:
                if d > mx[x]:
                    mx[x] = d
                if d < mn[x]:
                    mn[x] = d
    norm = [[0.0] * ny for _ in range(nx)]
    for x in range(nx):  # pass 2 :293-300 (0/0 = NaN, x/0 = Inf as in Java)
        for y in range(ny):
            if dist[x][y] >= 0:
                num, den = dist[x][y] - mn[x], mx[x] - mn[x]
                norm[x][y] = (num / den) if den != 0 else (math.nan if num == 0 or num != num else math.copysign(math.inf, num))
            else:
                norm[x][y] = dist[x][y]
    scores = []
    for y, t in enumerate(taxa):  # pass 3 :305-326
        total, count = 0.0, 0
        for x, c in enumerate(columns):
            if c == selected_col:
                continue
            if dist[x][y] >= 0:
                total += norm[x][y]
                count += 1
        total = (total / count) if count else math.nan  # 0.0 / 0
        scores.append((t, total))

    def jlong(v):  # (long) (d * 100000): NaN -> 0, saturating
        if v != v:
            return 0
        return make_long(v)

    def cmp(s1, s2):  # Score.compareTo :106-127
        if s1[0] == ref_taxon:
            return -1
        if s2[0] == ref_taxon:
            return +1
        return _java_long_to_int(jlong(s1[1]) - jlong(s2[1]))

    java_sort(scores, cmp)  # :329
    order = [taxa.index(t) for (t, _) in scores]  # sequencesList.indexOf(seqName): the first taxon of that name
    dist2 = [[dist[x][o] for o in order] for x in range(nx)]
    norm2 = [[norm[x][o] for o in order] for x in range(nx)]
    ranks = [[0] * ny for _ in range(nx)]
    for x in range(nx):  # :371-385
        ll = sorted(v for v in dist2[x] if v >= 0)
        for y in range(ny):
            ranks[x][y] = ll.index(dist2[x][y]) if dist2[x][y] >= 0 else int(math.floor(dist2[x][y]))
    return dict(columns=columns, taxa=[t for (t, _) in scores], distances=dist2, norm_distances=norm2, scores=[s for (_, s) in scores], ranks=ranks)
