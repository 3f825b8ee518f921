#!/bin/bash
# usage: tools/sass_of.sh <lib.so> <mangled-name-substring>   -> "addr  instruction" lines of that kernel
cuobjdump -sass "$1" | awk -v pat="$2" '/Function :/{on = index($0, pat) > 0} on' | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -E 's/^\s+\/\*([0-9a-f]+)\*\/\s+/\1 /; s/\s*\/\*.*$//'
