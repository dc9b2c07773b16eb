#!/bin/bash
cd "$GRAFT_REPO_ROOT"
echo "--- narrow (2 CTAs/SM)"; python tools/mlp_phases.py attention 2>&1 | tail -7
echo "--- wide (1 CTA/SM)"; RF_MLP_WIDE=1 python tools/mlp_phases.py attention 2>&1 | tail -7
