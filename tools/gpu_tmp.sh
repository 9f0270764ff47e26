#!/bin/bash
# scratch script for one-off gpurun calls; the committed state runs the smoke check
python -c "import __graft_entry__ as g; g.smoke()"
