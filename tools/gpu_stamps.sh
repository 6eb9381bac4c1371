python tools/stamps.py 2>&1 | cut -c1-900
echo COLD
python tools/stamps.py --cold 2>&1 | cut -c1-900
