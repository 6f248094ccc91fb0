#!/bin/bash
# registers / stack (spill) bytes / static shared memory per kernel of a built library
cuobjdump --dump-resource-usage "$1" 2>/dev/null | awk '/Function/ {name=$2} /REG:/ {print name, $0}' | c++filt | \
  sed -E 's/kamr:://g; s/\(.*\): +/ /; s/CONSTANT.*//; s/LOCAL:0//' | grep -E "${2:-.}" | sort
