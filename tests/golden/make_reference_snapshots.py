"""Transcribe the reference's own insta snapshots (golden outputs of spenso's seeded tests)
into tests/golden/reference_snapshots.json.

Run HERE only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_reference_snapshots.py

Source files: /root/reference/spenso/src/snapshots/spenso__tests__*.snap, written by the tests in
spenso/src/tests.rs (rng_is_deterministic :150, single_fiber :174, fibers :229,
fiber_from_structure :244, trace :295-306, scalar_and_dim1_conract :385-414, single_contract :534-583,
multi_contract :744-797, arithmetic_data :1143-1156).  Only numbers and slot descriptions are
taken — no reference code.  The inputs of those tests come from a seeded Rust RNG that
tests/_rustrand.py restates; tests/test_reference_snapshots.py regenerates the inputs, runs the
oracle / the device path on them and compares against this file.
"""
import json
import os
import re
import sys

SNAP = "/root/reference/spenso/src/snapshots"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_snapshots.json")


def body(name):
    t = open(os.path.join(SNAP, name)).read()
    return t.split("---", 2)[2]


def ints(text):
    return [int(x) for x in re.findall(r"-?\d+", text)]


def ron_slots(text):
    """Slot(aind: Normal(a), rep: Representation(dim: Concrete(d), rep: Kind(id)))"""
    out = []
    for a, d, kind, rid in re.findall(
        r"aind:\s*Normal\((\d+)\),\s*rep:\s*Representation\(\s*dim:\s*Concrete\((\d+)\),\s*rep:\s*(\w+)\((-?\d+)\)", text
    ):
        out.append({"aind": int(a), "dim": int(d), "kind": kind, "id": int(rid)})
    return out


def flat_lists(text):
    """split a RON tuple '(Struct(...), [FlatIndex..], [FlatIndex..])' into its index lists"""
    lists = []
    for m in re.finditer(r"\[\s*((?:FlatIndex\(\s*index:\s*\d+,\s*\),\s*)+)\]", text):
        lists.append([int(x) for x in re.findall(r"index:\s*(\d+)", m.group(1))])
    return lists


def main():
    if not os.path.isdir(SNAP):
        sys.exit("reference snapshots not found (run in the build container)")
    g = {}
    g["rng_is_deterministic"] = ints(body("spenso__tests__rng_is_deterministic.snap"))
    g["trace_i8_seed3"] = 79  # tests.rs:306 assert_eq!(f.data(), vec![79])
    g["arithmetic_data"] = ints(body("spenso__tests__arithmetic_data.snap"))
    g["scalar_and_dim1"] = {
        "common": body("spenso__tests__scalar_and_dim1_conract.snap").split(),
        "structa": body("spenso__tests__scalar_and_dim1_conract-2.snap").split(),
        "structb": body("spenso__tests__scalar_and_dim1_conract-3.snap").split(),
        "result": body("spenso__tests__scalar_and_dim1_conract-4.snap").split(),
        "data": ints(body("spenso__tests__scalar_and_dim1_conract-5.snap")),
    }
    sc = body("spenso__tests__single_contract.snap")
    a_txt, b_txt = sc.split("), VecStructure(")
    g["single_contract"] = {
        "structa": ron_slots(a_txt),
        "structb": ron_slots(b_txt),
        "result": ron_slots(body("spenso__tests__single_contract-2.snap")),
    }
    mc = body("spenso__tests__multi_contract.snap")
    data_txt, st_txt = mc.split("structure: VecStructure(")
    g["multi_contract"] = {"data": ints(data_txt), "result": ron_slots(st_txt)}
    fb = body("spenso__tests__fibers.snap")
    fl = flat_lists(fb)
    g["fibers"] = {"structure": ron_slots(fb), "fiter": fl[0], "fciter": fl[1]}
    sf = body("spenso__tests__single_fiber.snap")
    fl = flat_lists(sf)
    g["single_fiber"] = {"structure": ron_slots(sf), "fciter": fl[0], "fiberclass": fl[1]}
    ff = body("spenso__tests__fiber_from_structure.snap")
    # YAML: a 2-tuple of lists of {indices: [5 ints]}; the second list starts at the second "- - indices"
    first, second = re.split(r"\n- - indices:", ff)[1:3]
    def groups(t):
        v = ints(t)
        return [v[i : i + 5] for i in range(0, len(v), 5)]
    g["fiber_from_structure"] = {"fiberclass_iter": groups(first), "fiber_iter": groups(second)}
    with open(OUT, "w") as f:
        json.dump(g, f, separators=(",", ":"))
    print({k: (len(v) if hasattr(v, "__len__") else v) for k, v in g.items()})
    print("fibers", len(g["fibers"]["fiter"]), len(g["fibers"]["fciter"]),
          "single_fiber", len(g["single_fiber"]["fciter"]), len(g["single_fiber"]["fiberclass"]),
          "ffs", len(g["fiber_from_structure"]["fiberclass_iter"]), len(g["fiber_from_structure"]["fiber_iter"]))


if __name__ == "__main__":
    main()
