cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "small_dense or bruss" 2>&1 | tail -3
