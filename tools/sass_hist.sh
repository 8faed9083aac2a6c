#!/bin/bash
# instruction histogram of one kernel of the built library:  tools/sass_hist.sh <mangled-name> [so]
SO=${2:-superslomo-videointerpolation-pytorch_b200/libssm_b200.so}
cuobjdump -sass -fun "$1" "$SO" | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]{4}\*\/\s+//' | sed -E 's/^@!?U?P[0-9T]+\s+//' | awk '{print $1}' | sed 's/\..*//;s/;//' | sort | uniq -c | sort -rn
