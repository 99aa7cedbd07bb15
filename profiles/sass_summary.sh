#!/bin/bash
# SASS evidence of the Blackwell-native instructions in the shipped library (run on the CPU box; cuobjdump needs no GPU).
# usage: bash profiles/sass_summary.sh > profiles/sass_summary.txt
LIB=transformer-mm-explainability_b200/libmmx.so
echo "library: $LIB   sources hash: $(python -c 'import importlib.util as u; s=u.spec_from_file_location("b","bench.py"); m=u.module_from_spec(s); s.loader.exec_module(m); print(m.csrc_hash())')"
cuobjdump -sass $LIB > /tmp/mmx.sass
for m in UTCHMMA UTCQMMA UTMALDG UTMASTG UBLKCP LDTM STTM UTCBAR "SYNCS" HMMA LDSM LDGSTS F2FP; do
  printf "%-10s %6d\n" "$m" "$(grep -c -E "(^|[^A-Z])$m" /tmp/mmx.sass)"
done
echo
echo "per kernel family (instantiations summed): UTCHMMA / UTMALDG / LDTM / STTM / HMMA / LDSM / LDGSTS / F2FP"
python3 - <<'PY'
import re, collections
agg = collections.defaultdict(lambda: [0] * 8)
keys = ["UTCHMMA", "UTMALDG", "LDTM", "STTM", "HMMA", "LDSM", "LDGSTS", "F2FP"]
name = None
for line in open("/tmp/mmx.sass"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        raw = m.group(1)
        k = re.search(r"(gemm_\w+_kernel|attention_\w+_kernel|rule6_chain_kernel|avg_heads\w*_kernel|bmm_add_kernel|split_planes_kernel)", raw)
        name = k.group(1) if k else None
        continue
    if name:
        for i, key in enumerate(keys):
            if re.search(r"(^|[^A-Z])" + key, line):
                agg[name][i] += 1
for n, v in sorted(agg.items()):
    print(f"{n:32s} " + " / ".join(str(x) for x in v))
PY
