"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): python tests/sanitize_smoke.py case1,case2"""
import sys, numpy as np
import os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
import oracle
from rx_tools_b200 import fm, power, synth
from cases import fm_cases, fm_optional_cases
port=oracle.port()
names=sys.argv[1].split(',')
for c in fm_cases()+fm_optional_cases():
    if c.name not in names: continue
    x=c.make_input()[:2*131072]
    d=fm.FmDemod(c.params); got=d.full_demod(x,2*32768); want=port.fm_run(c.params,x,2*32768)
    print(c.name, np.array_equal(got,want) if c.exact else np.abs(got.astype(int)-want).max()); d.close()
plan=power.plan_range("100M:101M:1k"); win=power.window_table("hamming",1024)
hb=synth.power_hops(3,1,plan.buf_len,seed=1); sc=power.PowerScanner(plan,win); sc.scanner(hb,3); a,s=sc.read()
wa,ws=port.power_scan(oracle.PowerParams(bin_e=10,buf_len=plan.buf_len),win,hb,3,1); print("power",np.array_equal(a,wa)); sc.close()
plan=power.plan_range("100M:100.1M:100"); win=power.window_table("bartlett",1<<plan.bin_e)
hb=synth.power_hops(2,1,plan.buf_len,seed=2); sc=power.PowerScanner(plan,win); sc.scanner(hb,2); a,s=sc.read()
wa,ws=port.power_scan(oracle.PowerParams(bin_e=plan.bin_e,buf_len=plan.buf_len,downsample=plan.downsample),win,hb,2,1); print("power ds",np.array_equal(a,wa)); sc.close()
