#!/bin/bash
cd "$(dirname "$0")/.." && python -c "
import sys; sys.path.insert(0,'.')
import __graft_entry__ as g; g.build()" 2>&1 | tail -${1:-3}
