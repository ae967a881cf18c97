#!/bin/bash
# usage: tools/sass_hist.sh <kernel-name-substring>  — opcode histogram of one kernel of libalego_b200.so
SO=/root/repo/a-lego-loam_b200/csrc/libalego_b200.so
cuobjdump -sass $SO | awk -v pat="$1" '
/Function :/ { on = index($0, pat) > 0 }
on && /^ +\/\*[0-9a-f]+\*\/ +[A-Z@!]/ { op=$2; if (op ~ /^@/) op=$3; sub(/\..*/, "", op); sub(/;/, "", op); c[op]++; n++ }
END { for (k in c) printf "%6d %s\n", c[k], k | "sort -rn"; close("sort -rn"); print n, "instructions total" }'
