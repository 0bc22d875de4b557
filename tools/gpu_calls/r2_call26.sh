#!/bin/bash
for mode in 0 1 2; do
TAG=mode$mode HYPREC_CHOL_FINALIZE=$mode python tools/debug_citeulike_t.py 2>&1 | grep "rep 0" | grep "128..191"
done
