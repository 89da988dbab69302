#!/bin/bash
# bring the private work tree (.wip/w, branch wip) into the main tree: fast-forward the sources, copy the built libraries
set -e
cd /root/repo/.wip/w && git add -A && (git diff --cached --quiet || git commit -q -m "${1:-wip}")
cd /root/repo && git merge -q --ff-only wip
cp .wip/w/mimsem_b200/*.so mimsem_b200/
rsync -a .wip/w/mimsem_b200/csrc/build/ mimsem_b200/csrc/build/ 2>/dev/null || cp -r .wip/w/mimsem_b200/csrc/build mimsem_b200/csrc/
rsync -a .wip/w/mimsem_b200/host/build/ mimsem_b200/host/build/ 2>/dev/null || true
git log --oneline | head -1
