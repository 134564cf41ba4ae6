#!/bin/bash
# usage: tools/sass_stats.sh obj.o <kernel-name-substring>   -> registers + instruction count + opcode histogram per matching kernel
obj=$1; pat=$2
cuobjdump -res-usage $obj 2>/dev/null | grep -A1 "Function.*$pat" | grep -E "Function|REG" | sed -e 's/Function \(.*\):/\1/' | paste - - | awk '{print $1, $2, $3, $5, $6}' | cut -c1-200
for f in $(cuobjdump -sass $obj | grep -E "Function : .*$pat" | awk '{print $3}'); do
  echo "== $f"
  cuobjdump -sass -fun "$f" $obj | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -e 's/^\s*\/\*[0-9a-f]*\*\/\s*//' | awk '{ if ($1 ~ /^@/) print $2; else print $1 }' | sed -e 's/\..*//' | sort | uniq -c | sort -rn | head -${3:-25} | awk '{printf "%s:%s ", $2, $1} END {print ""}'
  cuobjdump -sass -fun "$f" $obj | grep -cE "^\s+/\*[0-9a-f]{4}\*/"
done
