for cs in 4; do echo "CSPLIT=$cs"; DGB200_CSPLIT=$cs python tools/stamps.py 2>&1 | cut -c1-1200; done
