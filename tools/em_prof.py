import sys, json, time
sys.path.insert(0,'cellranger-atac_b200')
from cratac import _ffi
cases=json.load(open('tests/golden/em.json'))
ctx=_ffi.Context([("chr1",1000)])
for c in cases:
    if c.get('raises') or c['params'] is None: continue
    for fast in (2,):
        ctx.set_option('em_fast', fast)
        ctx.fit_threshold(c['values'], c['weights'], 0.2)
        t=time.time(); thr,p,st=ctx.fit_threshold(c['values'], c['weights'], 0.2); dt=time.time()-t
        dbg=ctx.em_debug(); tot=sum(dbg.values())
        print(c["name"], "fast", fast, "ms %.2f" % (dt*1e3), st, {k: round(v/max(st['em_iterations'],1)) for k,v in dbg.items()}, 'cyc/iter total', round(tot/max(st['em_iterations'],1)))
