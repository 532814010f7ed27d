"""tests/golden/make_decomposer_kats.py -- extract the reference's known-answer vectors for the decomposer (custom struct key)
overloads of cub::DeviceRadixSort into tests/golden/decomposer_kats.json.

Source: /root/reference/test/catch2_test_device_radix_sort_custom.cu:555-1690 -- `custom_t {float f; int unused; long long lli;}`
with `decomposer_t` returning (f, lli); every SECTION lists an input, optionally values and a bit range, and the expected
output.  Run in the build container (the reference tree is not on the GPU box); the JSON is committed.
    python tests/golden/make_decomposer_kats.py"""
import json
import os
import re

import numpy as np

SRC = "/root/reference/test/catch2_test_device_radix_sort_custom.cu"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "decomposer_kats.json")

text = open(SRC).read()
start = text.index("struct custom_t")
body = text[start:]


def parse_records(init: str):
    recs = []
    for m in re.finditer(r"\{\s*([^{},]+)\s*,\s*([^{}]+?)\s*\}", init):
        f_txt, i_txt = m.group(1).strip(), m.group(2).strip()
        f = float(f_txt.rstrip("f"))
        fbits = int(np.array([f], dtype=np.float32).view(np.uint32)[0])
        if f_txt.startswith("-") and f == 0.0:
            fbits = 0x80000000
        i = int(eval(re.sub(r"(\d+)ll\b", r"\1", i_txt)))  # e.g. `1ll << 61`
        recs.append([fbits, i])
    return recs


cases = []
test_title = None
for m in re.finditer(r'CUB_TEST\("([^"]+)"|SECTION\("([^"]+)"\)\s*\{', body):
    if m.group(1):
        test_title = m.group(1)
        continue
    sec = m.group(2)
    # the section's text runs to the next SECTION / CUB_TEST
    nxt = re.search(r'SECTION\("|CUB_TEST\("', body[m.end():])
    chunk = body[m.end(): m.end() + (nxt.start() if nxt else len(body))]
    vecs = dict((name, init) for name, init in re.findall(
        r"thrust::device_vector<custom_t>\s+(\w+)\s*=\s*\{(.*?)\};", chunk, flags=re.S))
    ints = dict((name, [int(x) for x in init.replace("\n", " ").split(",") if x.strip()]) for name, init in re.findall(
        r"thrust::device_vector<int>\s+(\w+)\s*=\s*\{(.*?)\};", chunk, flags=re.S))
    kin = next((vecs[k] for k in ("in", "keys_in", "keys_buf") if k in vecs), None)
    kexp = next((vecs[k] for k in ("expected_output", "expected_keys") if k in vecs), None)
    if kin is None or kexp is None:
        continue
    bits = None
    mb = re.search(r"begin_bit\s*=\s*sizeof\(long long int\) \* 8 - (\d+)", chunk)
    me = re.search(r"end_bit\s*=\s*sizeof\(long long int\) \* 8 \+ (\d+)", chunk)
    if mb and me:
        bits = [64 - int(mb.group(1)), 64 + int(me.group(1))]
    vin = next((ints[k] for k in ("vals_in", "vals_buf") if k in ints), None)
    vexp = ints.get("expected_vals")
    cases.append({"test": test_title, "section": sec, "descending": "Descending" in sec, "keys_in": parse_records(kin),
                  "keys_expected": parse_records(kexp), "values_in": vin, "values_expected": vexp, "bits": bits,
                  "fields": ["f32", "i64"]})

json.dump({"source": SRC + ":555-1690", "cases": cases}, open(OUT, "w"), indent=1)
print(len(cases), "cases ->", OUT)
for c in cases:
    print(" ", c["test"][-28:], "|", c["section"], "| n =", len(c["keys_in"]), "| bits", c["bits"], "| values", c["values_in"] is not None)
