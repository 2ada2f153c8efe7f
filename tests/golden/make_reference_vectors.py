#!/usr/bin/env python3
"""Extract the golden vectors of the reference's in-file unit tests.

Reads /root/reference/src/lib.rs (only available in the build container, never
on the GPU box) and writes tests/golden/reference_unit_vectors.json.  The JSON
holds DATA only -- the array literals and expected answers of the seven
#[test] functions that pin the hot path:

  test_find_px                lib.rs:259-291
  test_merge_eq               lib.rs:308-311
  test_merge_ord_small_big    lib.rs:336-344
  test_merge_ord_big_small    lib.rs:369-377
  test_find_merge             lib.rs:447-465
  test_make_colour_map        lib.rs:544-587
  test_recolour               lib.rs:594-626

Run:  python tests/golden/make_reference_vectors.py
"""
import json
import os
import re
import sys

REF = "/root/reference/src/lib.rs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_unit_vectors.json")


def body(lines, name):
    """Source text of `fn name() { ... }` plus its 1-based (start, end) lines."""
    start = next(i for i, l in enumerate(lines) if re.match(rf"\s*fn {name}\(\)", l))
    depth, i = 0, start
    while True:
        depth += lines[i].count("{") - lines[i].count("}")
        if depth == 0 and "{" in "".join(lines[start:i + 1]):
            break
        i += 1
    return "".join(lines[start:i + 1]), (start + 1, i + 1)


def arrays(text):
    """All nd::array![[..],[..]] literals in order, as nested int lists."""
    out = []
    for m in re.finditer(r"nd::array!\[(.*?)\];", text, re.S):
        rows = re.findall(r"\[([0-9,\s]+)\]", m.group(1))
        out.append([[int(v) for v in r.replace(" ", "").strip(",").split(",")] for r in rows])
    return out


def int_lists(s):
    return [[int(v) for v in grp.split(",")] for grp in re.findall(r"\[([0-9,\s]+)\]", s)]


def main():
    if not os.path.exists(REF):
        sys.exit(f"{REF} not found: run this in the build container")
    lines = open(REF, encoding="utf-8").read().splitlines(keepends=True)
    vec = {"source": "smups/rustronomy-watershed v0.4.1 src/lib.rs"}

    t, ln = body(lines, "test_find_px")
    a = arrays(t)
    vec["test_find_px"] = {
        "lines": ln, "input": a[0], "colours": a[1],
        "lvl": int(re.search(r"colours\.view\(\), (\d+)\)", t).group(1)),
        "must_contain": [list(map(int, p)) for p in
                         re.findall(r"\((\d+), (\d+)\)", re.search(r"answer1 = \[(.*?)\];", t).group(1))],
    }

    t, ln = body(lines, "test_merge_eq")
    m = re.search(r"Merge\(\[(\d+), (\d+)\]\), Merge\(\[(\d+), (\d+)\]\)", t)
    vec["test_merge_eq"] = {"lines": ln, "equal": [[int(m.group(1)), int(m.group(2))],
                                                    [int(m.group(3)), int(m.group(4))]]}

    for name in ("test_merge_ord_small_big", "test_merge_ord_big_small"):
        t, ln = body(lines, name)
        cases = re.findall(
            r"cmp\(&Merge\(\[(\d+), (\d+)\]\), &Merge\(\[(\d+), (\d+)\]\)\), (\w+)\)", t)
        vec[name] = {"lines": ln, "cases": [
            {"a": [int(c[0]), int(c[1])], "b": [int(c[2]), int(c[3])], "ordering": c[4]} for c in cases]}

    t, ln = body(lines, "test_find_merge")
    vec["test_find_merge"] = {
        "lines": ln, "input": arrays(t)[0],
        "answer": int_lists(re.search(r"answer = vec!\[(.*?)\];", t).group(1)),
    }

    t, ln = body(lines, "test_make_colour_map")
    scen = []
    cur = None
    for raw in t.splitlines():
        s = raw.strip()
        if s.startswith("cmap = ["):
            cur = {"start": int_lists(s)[0], "steps": []}
            scen.append(cur)
        elif "vec![Merge" in s:
            cur["steps"].append(int_lists(s[s.index("vec!"):]))
        elif s.startswith("assert!(cmap =="):
            cur["expect"] = int_lists(s)[0]
    vec["test_make_colour_map"] = {"lines": ln, "shuffles": 10, "scenarios": scen}

    t, ln = body(lines, "test_recolour")
    a = arrays(t)
    cm = [int_lists(m)[0] for m in re.findall(r"let cmap = (\[.*?\]);", t)]
    vec["test_recolour"] = {"lines": ln, "input": a[0], "answer": a[1], "cmap": cm[0], "stale_cmap": cm[1]}

    with open(OUT, "w") as f:
        json.dump(vec, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
