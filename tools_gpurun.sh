#!/bin/bash
# usage: tools_gpurun.sh <timeout> '<command>'  — rebuilds the library first so the snapshot is never stale
set -e
cd /root/repo
python -c "import importlib; p = importlib.import_module('nceplibs-ip_b200'); p.build_library()" 
make -s -C oracle
exec /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
