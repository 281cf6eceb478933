export DBG_SKIP_ALONE=1 DBG_REPS=8
echo "== merged (after fix)"; python tools/debug_group.py 10 2>&1 | tail -4
unset DBG_SKIP_ALONE; export DBG_REPS=2
echo "== vs alone"; python tools/debug_group.py 10 2>&1 | grep -v "\[\], ML \[\]" | tail -4
