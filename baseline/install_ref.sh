#!/bin/sh
# Copy the reference's pure-Python hot path next to the bench so that it travels to the GPU box (baseline/_ref is
# git-ignored, not gpurun-ignored).  `pip install --target baseline/_ref /root/reference` would do the same thing:
# the reference's setup.py is `find_packages()` over spatten_llm/ and nothing else (setup.py:1-6).
set -e
here=$(cd "$(dirname "$0")" && pwd)
mkdir -p "$here/_ref"
rm -rf "$here/_ref/spatten_llm"
cp -r "${1:-/root/reference}/spatten_llm" "$here/_ref/"
echo "installed $(ls "$here/_ref/spatten_llm" | wc -l) files into $here/_ref/spatten_llm"
