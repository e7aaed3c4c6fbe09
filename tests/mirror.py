"""An independent restatement of the reference's hot path, importable by the tests: a line-by-line Python mirror of the
Rust code that drives Python's `re` with the *same regex string* the reference builds (info.rs:263-298), keeps
`fix_error`'s loop/flag structure (parse.rs:553-593) and the f32 arithmetic of `low_quality` (parse.rs:331-375,
numpy.float32).  Test infrastructure: tests/golden/make_golden.py writes the committed fixtures with it and
tests/test_mirror_fuzz.py runs it against the C++ oracle (oracle/) on hundreds of random schemes — two restatements
written separately that have to agree read by read.  Citations are file:line into the reference repository.
"""
import random
import re

import numpy as np


# ----------------------------------------------------------------------------- info.rs:215-310
class SequenceFormat:
    def __init__(self, text):
        data = "".join(l for l in text.splitlines() if not l.startswith("#"))
        self.format_string = ""
        self.regions_string = ""
        self.constant_region_length = 0
        self.barcode_num = 0
        self.barcode_lengths = []
        self.sample_length_option = None
        self.random_barcode = False
        self.sample_barcode = False
        regex_string = ""
        barcode_search = re.compile(r"(?i)(\{\d+\})|(\[\d+\])|(\(\d+\))|N+|[ATGC]+")
        for group in barcode_search.finditer(data):
            group_str = group.group(0)
            group_name = None
            if "[" in group_str:
                group_name = "sample"
                self.sample_barcode = True
            elif "{" in group_str:
                self.barcode_num += 1
                group_name = "barcode%d" % self.barcode_num
            elif "(" in group_str:
                group_name = "random"
                self.random_barcode = True
            if group_name is not None:
                digits = int(re.search(r"\d+", group_str).group(0))
                regex_string += "(?P<%s>.{%d})" % (group_name, digits)
                if group_name == "sample":
                    self.sample_length_option = digits
                    push_char = "S"
                elif "barcode" in group_name:
                    self.barcode_lengths.append(digits)
                    push_char = "B"
                else:
                    push_char = "R"
                self.regions_string += push_char * digits
                self.format_string += "N" * digits
            elif "N" in group_str:
                regex_string += "[AGCT]{%d}" % group_str.count("N")
                self.format_string += group_str
            else:
                regex_string += group_str.upper()
                self.format_string += group_str
                self.regions_string += "C" * len(group_str)
                self.constant_region_length += len(group_str)
        self.length = len(self.format_string)
        self.regex_string = regex_string
        self.format_regex = re.compile(regex_string)


# ----------------------------------------------------------------------------- info.rs:490-543
def max_seq_errors(sample_errors, sample_size, barcode_errors, barcode_sizes, constant_errors, constant_size):
    if sample_size is not None:
        max_sample = sample_errors if sample_errors is not None else sample_size // 5
    else:
        max_sample = 0
    max_barcode = [barcode_errors if barcode_errors is not None else s // 5 for s in barcode_sizes]
    max_constant = constant_errors if constant_errors is not None else constant_size // 5
    return max_constant, max_sample, max_barcode


# ----------------------------------------------------------------------------- info.rs:364-456
def rust_lines(text):
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    return [l[:-1] if l.endswith("\r") else l for l in lines]


def sample_conversion(text):
    out = {}
    for line in rust_lines(text)[1:]:
        f = line.split(",")
        if len(f) >= 2:
            out[f[0]] = f[1]
        else:
            out[""] = ""
    return out


def barcode_conversion(text, barcode_num):
    out = [dict() for _ in range(barcode_num)]
    for line in rust_lines(text)[1:]:
        f = line.split(",")
        barcode, bid, num = (f[0], f[1], f[2]) if len(f) >= 3 else ("", "", "")
        out[int(num) - 1][barcode] = bid
    assert all(len(h) for h in out)
    return out


# ----------------------------------------------------------------------------- parse.rs:553-593
def fix_error(mismatch_seq, possible_seqs, mismatches):
    best_match = None
    best_mismatch_count = mismatches + 1
    keep = True
    for true_seq in possible_seqs:
        mm = 0
        for possible_char, current_char in zip(true_seq, mismatch_seq):
            if possible_char != current_char and current_char != "N" and possible_char != "N":
                mm += 1
            if mm > best_mismatch_count:
                break
        if mm == best_mismatch_count:
            keep = False
        if mm < best_mismatch_count:
            keep = True
            best_mismatch_count = mm
            best_match = true_seq
    return best_match if keep and best_match is not None else None


class Decoder:
    """parse.rs:15-164 (SequenceParser) + info.rs:661-809 (Results) for one thread."""

    def __init__(self, fmt, samples_hash, counted_hash, max_errors, min_quality):
        self.fmt = fmt
        self.samples_hash = samples_hash
        self.sample_seqs = list(samples_hash.keys())
        self.counted_hash = counted_hash
        self.counted_seqs = [list(h.keys()) for h in counted_hash]
        self.max_constant, self.max_sample, self.max_barcode = max_errors
        self.min_quality = np.float32(min_quality)
        self.counters = dict(matched=0, constant_region=0, sample_barcode=0, barcode=0, duplicates=0, low_quality=0)
        # Results::new (info.rs:678-732)
        self.random_mode = fmt.random_barcode
        self.table = {}
        self.sample_conversion_omited = False
        if samples_hash:
            for s in samples_hash:
                self.table[s] = {}
        elif not fmt.sample_barcode:
            self.table["barcode"] = {}
        else:
            self.sample_conversion_omited = True

    # parse.rs:287-313 + 270-283
    def fix_constant_region(self, sequence):
        fs = self.fmt.format_string
        length_diff = len(sequence) - len(fs)
        possible = [sequence[i:i + len(fs)] for i in range(max(length_diff, 0))]
        best = fix_error(fs, possible, self.max_constant)
        if best is None:
            return "", -1
        # the winner is unique, so its index is well defined
        idx = [i for i, w in enumerate(possible) if w == best]
        fixed = "".join(o if n == "N" else n for o, n in zip(best, fs))
        return fixed, idx[0]

    # parse.rs:323-375
    def low_quality(self, quality, start):
        scores = []
        previous_type = "\0"
        qs = [(ord(ch) - 33) & 0xFF for ch in quality]
        for score, seq_type in zip(qs[start:], self.fmt.regions_string):
            if seq_type != previous_type:
                if scores:
                    total = np.float32(0)
                    for s in scores:
                        total = np.float32(total + np.float32(s))
                    average = np.float32(total / np.float32(len(scores)))
                    if average < self.min_quality:
                        return True
                    scores = []
                previous_type = seq_type
                if seq_type != "C":
                    scores = [score]
            elif seq_type != "C":
                scores.append(score)
        return False

    # info.rs:735-808
    def add_count(self, sample, random_barcode, barcode_string):
        if self.sample_conversion_omited and sample not in self.table:
            self.table[sample] = {}
        if not self.random_mode:
            if sample in self.table:
                self.table[sample][barcode_string] = self.table[sample].get(barcode_string, 0) + 1
            return True
        key = "barcode" if sample == "" else sample
        if key in self.table:
            h = self.table[key]
            if barcode_string not in h:
                h[barcode_string] = {random_barcode or ""}
            else:
                if (random_barcode or "") in h[barcode_string]:
                    return False
                h[barcode_string].add(random_barcode or "")
                return True
        else:
            self.table[sample] = {barcode_string: {random_barcode or ""}}
        return True

    # parse.rs:89-148 + 439-524 + 55-70
    def process(self, sequence, quality):
        out = dict(status="constant_region", offset=-1, repaired=False, sample="", barcodes="", random=None)
        repaired_at = None
        if not self.fmt.format_regex.search(sequence):
            sequence, idx = self.fix_constant_region(sequence)
            if idx >= 0:
                repaired_at = idx
        m = self.fmt.format_regex.search(sequence)
        if m is None:
            self.counters["constant_region"] += 1
            return out
        out["offset"] = repaired_at if repaired_at is not None else m.start()
        out["repaired"] = repaired_at is not None
        if self.min_quality > 0.0:
            if self.low_quality(quality, m.start()):
                self.counters["low_quality"] += 1
                out["status"] = "low_quality"
                return out
        groups = m.groupdict()
        if "sample" in groups:
            s = groups["sample"]
            if not self.sample_seqs or s in self.samples_hash:
                sample = s
            else:
                sample = fix_error(s, self.sample_seqs, self.max_sample)
                if sample is None:
                    self.counters["sample_barcode"] += 1
                    out["status"] = "sample_barcode"
                    return out
        else:
            sample = "barcode"
        counted = []
        for index in range(self.fmt.barcode_num):
            b = groups["barcode%d" % (index + 1)]
            if self.counted_seqs and b not in self.counted_hash[index]:
                b = fix_error(b, self.counted_seqs[index], self.max_barcode[index])
                if b is None:
                    self.counters["barcode"] += 1
                    out["status"] = "barcode"
                    out["sample"] = sample
                    return out
            counted.append(b)
        rnd = groups.get("random")
        out.update(sample=sample, barcodes=",".join(counted), random=rnd)
        if self.add_count(sample, rnd, ",".join(counted)):
            self.counters["matched"] += 1
            out["status"] = "matched"
        else:
            self.counters["duplicates"] += 1
            out["status"] = "duplicate"
        return out

    # ------------------------------------------------------------------------- output.rs:74-485 (canonical form)
    def write(self, prefix, merge, enrich):
        fmt = self.fmt
        files = {}
        if enrich and fmt.barcode_num < 2:  # main.rs:22-25
            enrich = False
        sample_barcodes = list(self.table.keys())
        name = (lambda s: self.samples_hash.get(s, "barcode")) if self.samples_hash else (lambda s: s)
        if self.samples_hash:
            sample_barcodes.sort(key=name)
        else:
            sample_barcodes.sort()
        header = ",".join("Barcode_%d" % (i + 1) for i in range(fmt.barcode_num)) if fmt.barcode_num > 1 else "Barcode"
        if merge and len(sample_barcodes) == 1:
            merge = False

        def count_of(sample, code):
            v = self.table[sample].get(code)
            if v is None:
                return 0
            return len(v) if self.random_mode else v

        def convert(code):
            if not self.counted_hash:
                return code
            return ",".join(self.counted_hash[i][b] for i, b in enumerate(code.split(",")))

        single = {s: {} for s in sample_barcodes}
        double = {s: {} for s in sample_barcodes}
        written = set()
        merged_rows = []
        for s in sample_barcodes:
            rows = []
            for code in self.table[s]:
                c = count_of(s, code)
                w = convert(code)
                if merge and code not in written:
                    written.add(code)
                    merged_rows.append(w + "".join(",%d" % count_of(t, code) for t in sample_barcodes))
                rows.append("%s,%d" % (w, c))
                if enrich:
                    parts = w.split(",")
                    n = len(parts)
                    for i in range(n):  # info.rs:840-866
                        k = ",".join(parts[x] if x == i else "" for x in range(n))
                        single[s][k] = single[s].get(k, 0) + c
                    if fmt.barcode_num > 2:  # info.rs:869-904
                        for i in range(n - 1):
                            for j in range(i + 1, n):
                                k = ",".join(parts[x] if x in (i, j) else "" for x in range(n))
                                double[s][k] = double[s].get(k, 0) + c
            files["%s_%s_counts.csv" % (prefix, name(s))] = [header + ",Count"] + sorted(rows)
        if merge:
            files["%s_counts.all.csv" % prefix] = [header + "".join("," + name(s) for s in sample_barcodes)] + sorted(merged_rows)
        if enrich:
            kinds = [("Single", single)] + ([("Double", double)] if fmt.barcode_num > 2 else [])
            for desc, table in kinds:
                merged_rows = []
                for s in sample_barcodes:
                    rows = []
                    for code, c in table[s].items():
                        if merge and code not in written:
                            written.add(code)
                            merged_rows.append(code + "".join(",%d" % table[t].get(code, 0) for t in sample_barcodes))
                        rows.append("%s,%d" % (code, c))
                    files["%s_%s_counts.%s.csv" % (prefix, name(s), desc)] = [header + ",Count"] + sorted(rows)
                if merge:
                    files["%s_counts.all.%s.csv" % (prefix, desc)] = (
                        [header + "".join("," + name(s) for s in sample_barcodes)] + sorted(merged_rows))
        return files


# ----------------------------------------------------------------------------- read synthesis for the fixtures
BASES = "ACGT"


def rand_seq(rng, n):
    return "".join(rng.choice(BASES) for _ in range(n))


def mutate(rng, s, sub_rate, n_rate):
    out = []
    for ch in s:
        r = rng.random()
        if r < n_rate:
            out.append("N")
        elif r < n_rate + sub_rate:
            out.append(rng.choice([b for b in BASES if b != ch]))
        else:
            out.append(ch)
    return "".join(out)


def make_reads(rng, fmt, samples_hash, counted_hash, n_reads, read_len, sub_rate=0.02, n_rate=0.01, low_q_rate=0.15,
               sample_tail=None):
    """Reads that carry the scheme at a random offset, with substitutions / Ns / a few junk and edge reads."""
    reads = []
    fs = fmt.format_string
    L = fmt.length
    # positions of each capture group in the template
    m = re.compile(fmt.regex_string)
    spans = []
    pos = 0
    for tok in re.finditer(r"\(\?P<(\w+)>\.\{(\d+)\}\)|\[AGCT\]\{(\d+)\}|[A-Z]+", fmt.regex_string):
        if tok.group(1):
            spans.append((tok.group(1), pos, int(tok.group(2))))
            pos += int(tok.group(2))
        elif tok.group(3):
            pos += int(tok.group(3))
        else:
            pos += len(tok.group(0))
    assert pos == L, (pos, L)
    pool = []  # a small molecule pool so that UMI duplicates occur
    for i in range(n_reads):
        kind = rng.random()
        if kind < 0.06:
            seq = rand_seq(rng, read_len)  # junk
        else:
            if pool and rng.random() < 0.35:
                body = rng.choice(pool)
            else:
                body = list(fs)
                for j, ch in enumerate(body):
                    if ch == "N":
                        body[j] = rng.choice(BASES)
                for name, start, ln in spans:
                    if name == "sample" and samples_hash:
                        dna = rng.choice(list(samples_hash.keys()))
                    elif name.startswith("barcode") and counted_hash:
                        dna = rng.choice(list(counted_hash[int(name[7:]) - 1].keys()))
                    else:
                        continue
                    dna = (dna + rand_seq(rng, ln))[:ln]  # reference files may be longer/shorter than the slot (Q10)
                    body[start:start + ln] = list(dna)
                body = "".join(body)
                pool.append(body)
            max_off = read_len - L
            r = rng.random()
            if r < 0.08:
                off = max_off  # scheme ends exactly at the read end (Q3)
            elif r < 0.16:
                off = 0
            else:
                off = rng.randint(0, max_off)
            clean = rng.random() < 0.45
            body2 = body if clean else mutate(rng, body, sub_rate, n_rate)
            seq = rand_seq(rng, off) + body2 + rand_seq(rng, read_len - L - off)
            if rng.random() < 0.03:
                seq = mutate(rng, seq, 0.0, 0.05)
        if rng.random() < low_q_rate:
            qual = "".join(chr(33 + rng.randint(2, 30)) for _ in range(read_len))
        else:
            qual = "".join(chr(33 + rng.randint(18, 40)) for _ in range(read_len))
        reads.append((seq, qual))
    return reads
