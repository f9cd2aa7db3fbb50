cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -s -k "small_dense or jacobians" 2>&1 | grep -v "^$" | tail -40
