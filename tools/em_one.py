import sys, json
sys.path.insert(0,'cellranger-atac_b200')
from cratac import _ffi
cases=json.load(open('tests/golden/em.json'))
c=[x for x in cases if x['name']==(sys.argv[1] if len(sys.argv)>1 else 'GV4')][0]
ctx=_ffi.Context([("chr1",1000)])
ctx.set_option('em_fast', 2)
print(ctx.fit_threshold(c['values'], c['weights'], 0.2))
