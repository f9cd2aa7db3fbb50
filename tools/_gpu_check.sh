cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q -s -k "full_size_configs" 2>&1 | grep -v "^$" | tail -12
