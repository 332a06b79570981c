import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.quick_bench import run
run("cfg4", "auto", 4); run("cfg4", "auto", 4, variant="skew"); run("cfg4", "auto", 2)
run("cfg5", "auto", 4); run("cfg2", "auto", 1); run("cfg2", "auto", 2); run("cfg3", "auto", 1)
