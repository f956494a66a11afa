#!/bin/bash
# last check of the tree as it stands: smoke + the full GPU suite
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2ae_smoke.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2ae_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r2ae_pytest.txt
tail -n 2 gpurun_out/r2ae_smoke.txt; tail -n 3 gpurun_out/r2ae_pytest.txt
