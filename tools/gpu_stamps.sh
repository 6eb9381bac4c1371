python tools/stamps.py --big --cold 2>&1 | cut -c1-1200
