import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/fit-sne_b200')
import fitsne_b200 as fb, bench_util
N=1000000
row,col,val,labels = bench_util.knn_like_graph(N,15)
Y0 = bench_util.clustered_embedding(labels,2,170.0)
with fb.FitSNE(row,col,val,Y0) as t:
    for i in range(12):
        t0=time.time(); t.step(exaggeration=1.0,momentum=0.5,learning_rate=N/12.0,max_step_norm=5.0); t.synchronize(); print('step',i,'%.2f ms'%((time.time()-t0)*1e3), t.stats()['n_boxes'], file=sys.stderr)
