#!/bin/bash
python scripts/e2e_probe_mesh.py 2>&1 | grep -v "^Scene has"
