#!/bin/bash
# usage: tools/sass_mix.sh <lib-or-cubin> <function-name-regex> [divisor]
# prints the SASS opcode histogram of one kernel (optionally divided by `divisor`, e.g. butterflies per thread)
f=$1; pat=$2; div=${3:-1}
cuobjdump -sass "$f" 2>/dev/null | awk -v pat="$pat" '/Function :/{on=($0 ~ pat)} on' | grep -oE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+ )?[A-Z0-9_.]+" | awk '{print $NF}' | sort | uniq -c | sort -rn | awk -v d=$div '{if ($1/d >= 0.1) printf "%s:%.2f ", $2, $1/d; t+=$1} END{print "\nTOTAL =", t, " per unit =", t/d}'
