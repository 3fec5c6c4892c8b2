cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q -k test_wide_rows_longer_than_the_column_mean 2>&1 | grep -E "^E  .*Assert|passed|failed" | cut -c1-600 | head
