"""Compares tests/golden/vectors.json (derived from the oracle's two restatements) with the same schema dumped by the REAL
reference crate (rust/golden_dump.rs, run by a maintainer who has cargo).  Exit code 0 and "PINNED" = every value the CUDA path
and the oracle are tested against is the crate's own output; the oracle's "parity unpinned" header can then be dropped.

    python tests/golden/check_against_rust.py rust_golden.json
"""
import json
import os
import sys


def flatten(x, prefix=""):
    if isinstance(x, dict):
        for k in sorted(x):
            yield from flatten(x[k], prefix + "/" + str(k))
    elif isinstance(x, list):
        for i, v in enumerate(x):
            yield from flatten(v, prefix + "[%d]" % i)
    else:
        yield prefix, str(x)


def compare(ours, theirs):
    a, b = dict(flatten(ours)), dict(flatten(theirs))
    missing = sorted(set(a) - set(b))
    extra = sorted(set(b) - set(a))
    diff = sorted(k for k in set(a) & set(b) if a[k] != b[k])
    return missing, extra, diff


def main(argv):
    here = os.path.dirname(os.path.abspath(__file__))
    ours = json.load(open(os.path.join(here, "vectors.json")))
    text = open(argv[1]).read()
    if "GOLDEN " in text:  # raw cargo output is fine too
        text = [l for l in text.splitlines() if l.startswith("GOLDEN ")][-1][len("GOLDEN "):]
    theirs = json.loads(text)
    missing, extra, diff = compare(ours, theirs)
    for k in missing:
        print("not in the Rust dump:", k)
    for k in extra:
        print("only in the Rust dump:", k)
    for k in diff:
        print("DIFFERENT:", k)
    if missing or diff:
        print("NOT PINNED: %d differences, %d missing" % (len(diff), len(missing)))
        return 1
    print("PINNED: %d values equal the reference crate's output" % len(dict(flatten(ours))))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
