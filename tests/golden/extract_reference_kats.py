"""Regenerates tests/golden/reference_kats.json from the reference's own files (run in the build container, where
/root/reference exists; the GPU box only reads the committed JSON).  Every entry is a LITERAL value printed in the
reference's sources or doctests for the hot path -- no value is computed here.

    python tests/golden/extract_reference_kats.py [/root/reference]
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def lines(rel, lo, hi):
    with open(os.path.join(REF, rel), encoding="utf-8") as f:
        return f.read().split("\n")[lo - 1:hi]


def ints(s):
    return [int(v) for v in re.findall(r"-?\d+", s)]


kats = {}

# (1) src/cryptparams.jl:22-25 -- PALISADE (q, N, psi): "m => NegacyclicRing{GaloisField(q), N}(psi)"
rows = []
for ln in lines("src/cryptparams.jl", 22, 25):
    m = re.search(r"GaloisField\((\d+)\), (\d+)\}\((\d+)\)", ln)
    rows.append([int(m.group(1)), int(m.group(2)), int(m.group(3))])
kats["palisade_q_N_psi"] = {"source": "src/cryptparams.jl:22-25", "value": rows}

# (2) docs/src/man/background/rlwe.md:183-187 -- NegacyclicRing{F97, 4}() prints its minimal root
txt = "\n".join(lines("docs/src/man/background/rlwe.md", 180, 188))
m = re.search(r"NegacyclicRing\{𝔽₉₇,4\}\((\d+)\)", txt)
kats["minimal_root_q97_N4"] = {"source": "docs/src/man/background/rlwe.md:183-187", "value": int(m.group(1))}

# (3) docs/src/man/background/rlwe.md:190-212 -- operands and the products [p3*p4, p1^2, p1*p2]
ops = {}
for ln in lines("docs/src/man/background/rlwe.md", 193, 196):
    m = re.match(r"(p\d) = ℛ\(O\(\[([^\]]*)\]\)\)", ln)
    if m:
        ops[m.group(1)] = ints(m.group(2))
outs = [ints(ln) for ln in lines("docs/src/man/background/rlwe.md", 209, 211)]
kats["rlwe_products_q97"] = {"source": "docs/src/man/background/rlwe.md:190-212", "operands": ops,
                             "p3*p4": outs[0], "p1^2": outs[1], "p1*p2": outs[2]}

# (4) docs/src/man/encoding.md:14-24 -- GF(7), N = 2, psi = nothing (naive path): 3 * 4
out = [ints(ln)[0] for ln in lines("docs/src/man/encoding.md", 22, 23)]
kats["naive_product_q7_N2"] = {"source": "docs/src/man/encoding.md:14-24", "a": [3, 0], "b": [4, 0], "value": out}

# (5) docs/src/man/encoding.md:69-92 -- slots 1..10 times 10 at q = 65537, N = 2048
out = [ints(ln)[0] for ln in lines("docs/src/man/encoding.md", 81, 91)]
kats["slot_product_q65537_N2048"] = {"source": "docs/src/man/encoding.md:69-92", "a_slots_0_9": list(range(1, 11)), "b_slots_all": 10,
                                     "value_slots_0_10": out}

# (6) src/crt.jl:23-33 -- CRTEncoded{(5,7)}(3) * CRTExpand{11}: residues, and the Integer the docstring prints
txt = "\n".join(lines("src/crt.jl", 23, 33))
kats["crt_expand"] = {"source": "src/crt.jl:23-33", "x_residues": ints(re.search(r"\(\((\d+, \d+)\)\)", txt).group(1)),
                      "product_residues": ints(re.search(r"\(\((\d+, \d+, \d+)\)\)", txt).group(1)),
                      "docstring_integer": ints(lines("src/crt.jl", 32, 32)[0])[0],
                      "note": "the printed 333 contradicts the printed residues (333 mod 7 = 4); the residues (3,5,0) are the KAT"}

# (7) src/crt.jl:50-58 -- CRTResidual example: mod(3*invmod(77,5),5)*77
kats["crt_residual"] = {"source": "src/crt.jl:50-58", "value": ints(lines("src/crt.jl", 54, 54)[0])[0],
                        "formula_value": ints(lines("src/crt.jl", 57, 57)[0])[0], "basis_of_the_arithmetic": [5, 7, 11]}

# (8) test/bfv_crt.jl:43-46 and test/bfv_keyswitch.jl -- decrypt-level expectations (randomised keys)
kats["bfv_crt_test"] = {"source": "test/bfv_crt.jl:25-47", "plain_modulus": ints(lines("test/bfv_crt.jl", 27, 27)[0])[-1],
                        "plain0": ints(lines("test/bfv_crt.jl", 40, 40)[0])[-1], "square_hex": 0x24}

with open(os.path.join(HERE, "reference_kats.json"), "w", encoding="utf-8") as f:
    json.dump(kats, f, indent=1, ensure_ascii=False)
    f.write("\n")
print(json.dumps(kats, indent=1, ensure_ascii=False))
