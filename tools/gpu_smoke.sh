#!/bin/bash
# __graft_entry__.smoke() on the GPU box
python - <<'PY'
import __graft_entry__ as g
g.smoke()
PY
