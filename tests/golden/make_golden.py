#!/usr/bin/env python
"""Generate tests/golden/reference_vectors.json from the reference tree.

Run in the build container only (needs /root/reference, which does not exist on the
GPU box):  python tests/golden/make_golden.py

It extracts, without executing any reference code (Rust cannot be built here):
  * the toy FASTQ reads and the expected table     tests/test_data/toy.fq.gz, ground_truth.csv
  * the adapters and call used by the Python test  tests/test_vfind.py:5-13
  * AA_TABLE_CANONICAL and ASCII_TO_INDEX literals src/lib.rs:52-95
  * the four alignment accept/reject unit vectors  src/lib.rs:339-369 (+ constants :332-336)
  * the find_variants signature defaults           src/lib.rs:169-182
  * the threshold error text                       src/lib.rs:107-109
These are the only result-pinning fixtures the reference holds for the path.
"""
import csv
import gzip
import json
import os
import re
import sys

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")


def main():
    src = open(os.path.join(REF, "src/lib.rs")).read()

    # --- toy reads
    text = gzip.open(os.path.join(REF, "tests/test_data/toy.fq.gz"), "rt").read()
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    assert len(lines) % 4 == 0
    reads = []
    for i in range(0, len(lines), 4):
        assert lines[i].startswith("@") and lines[i + 2].startswith("+")
        reads.append({"header": lines[i][1:], "seq": lines[i + 1], "qual": lines[i + 3]})
    with open(os.path.join(REF, "tests/test_data/ground_truth.csv")) as f:
        rows = list(csv.reader(f))
    assert rows[0] == ["sequence", "count"]
    table = [[r[0], int(r[1])] for r in rows[1:]]

    # --- python test call
    pytest_src = open(os.path.join(REF, "tests/test_vfind.py")).read()
    m = re.search(r'\(\s*"([ACGT]+)"\s*,\s*"([ACGT]+)"\s*\)', pytest_src)
    adapters = [m.group(1), m.group(2)]

    # --- codon table: chars in declaration order == index c1*16 + c2*4 + c3
    block = src[src.index("static AA_TABLE_CANONICAL"):src.index("static ASCII_TO_INDEX")]
    block = block[block.index("= ["):]
    aa = re.findall(r"'(.)'", re.sub(r"//.*", "", block))
    assert len(aa) == 64, len(aa)
    block = src[src.index("static ASCII_TO_INDEX"):]
    block = block[block.index("= [") + 3:block.index("];")]
    idx = [int(x) for x in re.findall(r"\b\d+\b", re.sub(r"//.*", "", block))]
    assert len(idx) == 128, len(idx)

    # --- unit vectors
    consts = {k: float(v) if "." in v else int(v) for k, v in re.findall(
        r"const (\w+): (?:i32|f64) = (-?[\d.]+);", src)}
    units = []
    for name, adapter, seq, good in re.findall(
            r"fn (test_\w+_alignment)\(\) \{\s*let (?:prefix|suffix) = b\"([ACGT]+)\";\s*"
            r"let seq = b\"([ACGT]+)\";\s*let good_alignment = (true|false);", src):
        units.append({"name": name, "adapter": adapter, "seq": seq, "accept": good == "true",
                      "is_prefix": "prefix" in name})
    assert len(units) == 4, units

    # --- signature defaults
    sig = src[src.index("#[pyo3(signature = ("):]
    sig = sig[:sig.index("))]")]
    defaults = {}
    for k, v in re.findall(r"(\w+)=([^,\s]+)", sig):
        if v in ("true", "false"):
            defaults[k] = v == "true"
        elif "." in v:
            defaults[k] = float(v)
        else:
            defaults[k] = int(v)
    order = re.findall(r"^\s*(\w+)(?:=[^,]+)?,\s*$", sig, flags=re.M)

    err = re.search(r'"(Accept alignment threshold[^"]+)"', src).group(1)

    out = {
        "_generated_by": "tests/golden/make_golden.py from /root/reference (nsbuitrago/vfind)",
        "toy": {"reads": reads, "adapters": adapters, "table": table,
                "source": "tests/test_data/toy.fq.gz, tests/test_data/ground_truth.csv, tests/test_vfind.py:5-13"},
        "aa_table_canonical": "".join(aa),
        "ascii_to_index": idx,
        "unit_constants": consts,
        "unit_vectors": units,
        "signature_order": order,
        "signature_defaults": defaults,
        "threshold_error": err,
    }
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")
    print("wrote", OUT)


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; fixtures are committed, nothing to do")
    main()
