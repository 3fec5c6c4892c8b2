cd $GRAFT_REPO_ROOT
for sr in 2097152 1048576 1398144 699072 524288; do
echo "slice rows $sr"; STRSIM_B200_SLICE_ROWS=$sr python tools/plugin_e2e.py --pageable 2>/dev/null | cut -c1-330
done
