#!/bin/bash
# env-variable sweep of the XLong probe (no rebuild): each argument is one "VAR=value ..." set
for V in "$@"; do
  echo "== $V"; env $V timeout 120 python -m tests.probe_xlong 256 5 2>&1 | grep -E "B=256|rec_"
done
