cd $GRAFT_REPO_ROOT
nproc
for ct in 8 4 6 12 16; do
echo "copy threads $ct"; STRSIM_B200_COPY_THREADS=$ct python tools/plugin_e2e.py --pageable 2>/dev/null | cut -c1-200
done
