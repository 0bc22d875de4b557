#!/bin/bash
TAG=default python tools/debug_citeulike_t.py 2>&1 | grep -v "^$" | tail -10
TAG=nosmall HYPREC_SMALL_F64=0 python tools/debug_citeulike_t.py 2>&1 | tail -10
TAG=nomulti HYPREC_TC_MULTI=0 python tools/debug_citeulike_t.py 2>&1 | tail -10
