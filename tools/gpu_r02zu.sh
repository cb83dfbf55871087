#!/bin/bash
# config 5 per-clip times: tree of commit d4855af (before this session) vs the current tree, same box
export ORVB_NO_BUILD=1
echo "== old tree"; (cd _old_tree && timeout 300 python tools/time_clips.py 5 10 2>&1 | tail -2)
echo "== new tree"; timeout 300 python tools/time_clips.py 5 10 2>&1 | tail -2
echo "== old tree again"; (cd _old_tree && timeout 300 python tools/time_clips.py 5 6 2>&1 | tail -2)
