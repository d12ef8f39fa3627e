timeout 600 python tools/sweep_yield.py 4096 3 1,2,4 0,64,96,128,96:128,96:512 2>&1 | grep -v "^$"
