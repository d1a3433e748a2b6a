#!/bin/bash
# usage: tools/sass_fn.sh <lib.so> <mangled-name-substring>   -> compact SASS of the first matching function
cuobjdump -sass "$1" 2>&1 | awk -v pat="$2" '/Function :/{ if (f) exit; if (index($0, pat)) f=1 } f' \
  | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s*\/\* 0x[0-9a-f]+ \*\/$//'
