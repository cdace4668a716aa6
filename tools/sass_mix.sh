#!/bin/bash
# usage: tools/sass_mix.sh <lib.so|cubin> [kernel]   -- SASS opcode histogram per kernel
f=$1; k=${2:-}
cuobjdump -sass "$f" | awk -v want="$k" '
/Function :/ {fn=$3}
/^[ \t]+\/\*[0-9a-f]+\*\/[ \t]+/ { if (want=="" || fn==want) { op=$2; if (op ~ /^@/) op=$3; sub(/\..*/,"",op); sub(/;/,"",op); c[fn" "op]++; t[fn]++ } }
END { for (x in t) print t[x], x, "TOTAL"; for (x in c) print c[x], x }' | sort -k2,2 -k1,1nr
