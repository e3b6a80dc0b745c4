"""numpy model of the DEVICE algorithm (executable spec for fit-sne_b200/csrc) -- test infrastructure.

The CUDA path does not transliterate the reference's charge formulation {1,x,y,x^2+y^2}: in fp32 that loses
the 1e-4 gradient tolerance late in a run.  It uses the algebraically identical
"local offset" form instead: per point and node, b = y_point - X_node (computed from the in-box coordinate,
never from global coordinates); grids w1 = sum L, delta_k = sum L*b_k, wbb = sum L*|b|^2; kernels
Ksq=(1+r^2/df)^-(df+1), Kgrad_k = R_k*Ksq, Kb=(1+r^2/df)^-df on the node-offset lattice; outputs
v1 = Ksq*w1, B_k = Kgrad_k*w1 - Ksq*delta_k; force numerator sum_nodes L*(a_k*v1 + B_k); Z by Parseval in the
frequency domain.  ``device_gradient_rep(Y, p, ipi, min_int, df, ft)`` evaluates exactly that with dtype
``ft``; tests/test_device_model.py checks it against the reference's golden vectors (fp64: ~1e-12,
fp32: ~4e-7), which is the evidence that the reformulation is the same algorithm.
"""
import numpy as np
f32=np.float32
ALLOWED=[25,36,50,55,60,65,70,75,80,85,90,96,100,110,120,130,140,150,175,200]
def grid_params(Y, ipi, min_int):
    N,d = Y.shape
    if d==2:
        flat = Y.reshape(-1)
        # ascending prefix rule
        t=1
        while t < flat.size and flat[t] > flat[:t].max(): t+=1
        mn = flat[t:].min() if t<flat.size else np.inf
        mx = flat.max()
        B = int(max(min_int,(float(mx)-float(mn))/ipi))
        if B<200:
            B=[a for a in ALLOWED if a>=B][0]
    else:
        mn, mx = Y.min(), Y.max()
        B = int(max(min_int,(float(mx)-float(mn))/ipi))
    return float(mn), float(mx), B
def nice(n):
    while True:
        m=n
        for p in (2,3,5,7):
            while m%p==0: m//=p
        if m==1: return n
        n+=1
def device_gradient_rep(Y32, p, ipi, min_int, df, ft=np.float32):
    """returns F (=neg_f = F_rep/Z, shape N,d) and Z  using local-offset formulation; ft = float type for grids/FFT"""
    N,d = Y32.shape
    mn,mx,B = grid_params(Y32, ipi, min_int)
    G=p*B
    bw = (mx-mn)/B   # fp64
    # binning in fp64, mirror reference: idx = int((y-min)/bw) clamped; u = (y - (idx*bw+min))/bw
    y64 = Y32.astype(np.float64)
    idx = ((y64-mn)/bw).astype(np.int64)   # trunc toward zero
    idx = np.clip(idx,0,B-1)
    u = ((y64-(idx*bw+mn))/bw).astype(ft)      # in-box coord, fp32
    s = ((np.arange(p)+0.5)/p).astype(ft)
    den = np.array([np.prod([s[i]-s[j] for j in range(p) if j!=i]) for i in range(p)]).astype(ft)
    def lag(u):
        out = np.ones((N,p),ft)
        for j in range(p):
            for k in range(p):
                if k!=j: out[:,j]*= (u-s[k])
            out[:,j]/=den[j]
        return out
    bwf = ft(bw)
    M = nice(2*G)
    h = bw/p
    if d==2:
        Lx, Ly = lag(u[:,0]), lag(u[:,1])
        bx = bwf*(u[:,0:1]-s[None,:])   # N,p offsets  y - X_node
        by = bwf*(u[:,1:2]-s[None,:])
        W = np.zeros((4,G,G),ft)   # [row=y node][col=x node]
        node_x = idx[:,0:1]*p+np.arange(p)[None,:]
        node_y = idx[:,1:2]*p+np.arange(p)[None,:]
        for a in range(p):      # y
            for b in range(p):  # x
                L = Ly[:,a]*Lx[:,b]
                r,c = node_y[:,a], node_x[:,b]
                np.add.at(W[0],(r,c),L)
                np.add.at(W[1],(r,c),L*bx[:,b])
                np.add.at(W[2],(r,c),L*by[:,a])
                np.add.at(W[3],(r,c),L*(bx[:,b]**2+by[:,a]**2))
        # kernels
        dd = np.zeros(M); off = np.arange(M); sgn = np.where(off<G, off, np.where(off>M-G, off-M, 0)); valid = (off<G)|(off>M-G)
        Rx = (h*sgn)[None,:]*np.ones((M,1)); Ry=(h*sgn)[:,None]*np.ones((1,M)); V = valid[None,:]&valid[:,None]
        r2 = Rx**2+Ry**2
        Ksq = np.where(V,(1+r2/df)**(-(df+1)),0); Kb = np.where(V,(1+r2/df)**(-df),0)
        Kx = Rx*Ksq; Ky = Ry*Ksq
        F = lambda a: np.fft.rfft2(a.astype(ft)).astype(np.complex64 if ft==np.float32 else np.complex128)
        sc = ft(1.0/(M*M))
        Ksq_h, Kb_h, Kx_h, Ky_h = [F(k)*sc for k in (Ksq,Kb,Kx,Ky)]
        pad = np.zeros((4,M,M),ft); pad[:,:G,:G]=W
        Wh = [F(pad[t]) for t in range(4)]
        v1h = Ksq_h*Wh[0]; Bxh = Kx_h*Wh[0]-Ksq_h*Wh[1]; Byh = Ky_h*Wh[0]-Ksq_h*Wh[2]
        # Parseval Z: sum over full spectrum; weights for half spectrum
        wt = np.full(M//2+1,2.0); wt[0]=1.0
        if M%2==0: wt[-1]=1.0
        def dot(ah,bh): return float((np.real(np.conj(ah.astype(np.complex128))*bh.astype(np.complex128))*wt[None,:]).sum())
        if df==1.0:
            Z = dot(Wh[0],Kb_h*Wh[0]) + 2*dot(Wh[3],v1h) + 4*dot(Wh[1],Kx_h*Wh[0]) + 4*dot(Wh[2],Ky_h*Wh[0]) - 2*dot(Wh[1],Ksq_h*Wh[1]) - 2*dot(Wh[2],Ksq_h*Wh[2]) - N
        else:
            Z = dot(Wh[0],Kb_h*Wh[0]) - N
        inv = lambda a: np.fft.irfft2(a,s=(M,M)).astype(ft)*ft(M*M)   # irfft2 normalises by 1/M^2 ; cuFFT doesn't
        v1,Bx,By = inv(v1h),inv(Bxh),inv(Byh)
        Fx = np.zeros(N,ft); Fy=np.zeros(N,ft)
        for a in range(p):
            for b in range(p):
                L = Ly[:,a]*Lx[:,b]; r,c=node_y[:,a],node_x[:,b]
                Fx += L*(bx[:,b]*v1[r,c]+Bx[r,c]); Fy += L*(by[:,a]*v1[r,c]+By[r,c])
        Fo = np.stack([Fx,Fy],1)/ft(Z)
        return Fo, Z
    else:
        L = lag(u[:,0]); b = bwf*(u[:,0:1]-s[None,:]); node = idx[:,0:1]*p+np.arange(p)[None,:]
        W = np.zeros((3,G),ft)
        for a in range(p):
            np.add.at(W[0],node[:,a],L[:,a]); np.add.at(W[1],node[:,a],L[:,a]*b[:,a]); np.add.at(W[2],node[:,a],L[:,a]*b[:,a]**2)
        off=np.arange(M); sgn = np.where(off<G, off, np.where(off>M-G, off-M, 0)); valid=(off<G)|(off>M-G)
        R = h*sgn; r2=R**2
        Ksq = np.where(valid,(1+r2/df)**(-(df+1)),0); Kb=np.where(valid,(1+r2/df)**(-df),0); Kg = R*Ksq
        ct = np.complex64 if ft==np.float32 else np.complex128
        F = lambda a: np.fft.rfft(a.astype(ft)).astype(ct)
        sc=ft(1.0/M)
        Ksq_h,Kb_h,Kg_h = [F(k)*sc for k in (Ksq,Kb,Kg)]
        pad=np.zeros((3,M),ft); pad[:,:G]=W
        Wh=[F(pad[t]) for t in range(3)]
        v1h=Ksq_h*Wh[0]; Bh = Kg_h*Wh[0]-Ksq_h*Wh[1]
        wt=np.full(M//2+1,2.0); wt[0]=1.0
        if M%2==0: wt[-1]=1.0
        def dot(ah,bh): return float((np.real(np.conj(ah.astype(np.complex128))*bh.astype(np.complex128))*wt).sum())
        if df==1.0:
            Z = dot(Wh[0],Kb_h*Wh[0]) + 2*dot(Wh[2],v1h) + 4*dot(Wh[1],Kg_h*Wh[0]) - 2*dot(Wh[1],Ksq_h*Wh[1]) - N
        else:
            Z = dot(Wh[0],Kb_h*Wh[0]) - N
        inv = lambda a: np.fft.irfft(a,n=M).astype(ft)*ft(M)
        v1,Bv = inv(v1h),inv(Bh)
        Fo=np.zeros(N,ft)
        for a in range(p):
            Fo += L[:,a]*(b[:,a]*v1[node[:,a]]+Bv[node[:,a]])
        return (Fo/ft(Z))[:,None], Z


# ---------------------------------------------------------------------------------------------------------------
# Sharded zero-mean + bounds (k_shard_stats / k_center_shard in fitsne_kernels.cuh): per-rank records of the new,
# un-centred positions -> global mean and the bounds of the CENTRED embedding, including the reference's 2-D scan quirk
# (tsne.cpp:1045-1048), without any rank ever seeing the whole Y.  tests/test_shard_bounds_model.py checks this model
# against the literal scan on the centred array.
SHARD_HEAD = 8


def literal_bounds_scan(flat):
    """tsne.cpp:1045-1048 on the interleaved sequence: `if (v > max) max = v; else if (v < min) min = v;`"""
    mx, mn = -np.inf, np.inf
    for v in flat:
        if v > mx:
            mx = v
        elif v < mn:
            mn = v
    return mn, mx


def centre32(y32, mean64):
    """(float)((double) y - mean): the device's centring (monotonic in y)"""
    return (y32.astype(np.float64) - mean64).astype(np.float32)


def shard_records(Ynew32, world, head=SHARD_HEAD):
    """what k_shard_stats leaves on each rank: sums, per-dimension min (head points excluded on rank 0) / max, head values"""
    N, d = Ynew32.shape
    per = -(-N // world)
    recs = []
    for r in range(world):
        sl = Ynew32[r * per:min(N, (r + 1) * per)]
        nhead = min(head, len(sl)) if (r == 0 and d == 2) else 0
        rest = sl[nhead:]
        recs.append(dict(sum=sl.astype(np.float64).sum(0), mx=sl.max(0),
                         mn=rest.min(0) if len(rest) else np.full(d, np.inf, np.float32),
                         head=sl[:nhead].reshape(-1).copy()))
    return recs


def shard_bounds(recs, N, d):
    """what k_center_shard derives on every rank from the all-gathered records: (mean, bmin, bmax)"""
    mean = sum(r["sum"] for r in recs) / N
    bmn, bmx = np.float32(np.inf), np.float32(-np.inf)
    for r in recs:
        for k in range(d):
            bmn = min(bmn, centre32(np.float32(r["mn"][k]), mean[k]))
            bmx = max(bmx, centre32(np.float32(r["mx"][k]), mean[k]))
    run, ascending = -np.inf, True
    for i, v in enumerate(recs[0]["head"]):
        c = centre32(np.float32(v), mean[i & 1])
        if ascending and c > run:
            run = c
        else:
            ascending = False
            bmn = min(bmn, c)
    return mean, bmn, bmx
