import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.quick_bench import run
from pychain_b200 import engine
run("cfg4", "auto"); run("cfg2", "auto"); run("cfg3", "auto"); run("cfg5", "auto"); run("cfg5", "auto", Bover=256)
for e in (4, 8):
    os.environ["CHAIN_B200_PACK_EFFORT"] = str(e); engine._prepared_cache.clear(); print("effort", e); run("cfg4", "auto")
