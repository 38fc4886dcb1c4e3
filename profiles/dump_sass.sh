#!/bin/bash
# SASS of one kernel of spirit_b200/libSpirit.so + opcode counts (no GPU needed).
# usage: profiles/dump_sass.sh '<regex on the mangled name>' out.txt
LIB=${LIB:-spirit_b200/libSpirit.so}
cuobjdump -sass "$LIB" | awk -v re="$1" '
  /Function : / { on = ($0 ~ re) ; if (on) n++ }
  on && n == 1 { print }
' > "$2"
echo "# opcode counts" >> "$2"
grep -oE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+ )?[A-Z0-9_.]+" "$2" | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -40 >> "$2"
