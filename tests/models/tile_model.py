import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla, sys, time
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
from chm_b200 import synthetic
from oracle.pbsm3d_oracle import Config, PBSM3DOracle

def build(mesh, L=10):
    geo=mesh.geometry()
    o=PBSM3DOracle(Config.functional_test(L), mesh.neigh, geo, mesh.global_id, mesh.n_global, mesh.params)
    F=synthetic.forcing(geo.cx, geo.cy)
    asm=o.assemble(F,3600.0)
    return o, asm

def colour2(neigh):
    T=neigh.shape[0]; col=-np.ones(T,int)
    for s in range(T):
        if col[s]>=0: continue
        col[s]=0; st=[s]
        while st:
            i=st.pop()
            for n in neigh[i]:
                if n>=0 and col[n]<0: col[n]=1-col[i]; st.append(n)
    return col

def run(n, tile, ks):
    mesh=synthetic.uniform_mesh(n,n); T=mesh.n_local; L=10
    o,asm=build(mesh,L)
    col=colour2(mesh.neigh)
    # per-column tridiagonal solves: use dense batched Thomas
    diag,lat,below,above,rhs=asm.diag,asm.lat,asm.below,asm.above,asm.rhs
    def thomas(cols, g):  # g [L, len(cols)] -> solve T_col y = g
        d=diag[:,cols].copy(); lo=below[:,cols]; up=above[:,cols]; y=g.copy()
        cp=np.zeros_like(d)
        cp[0]=up[0]/d[0]; y[0]=y[0]/d[0]
        for z in range(1,L):
            den=d[z]-lo[z]*cp[z-1]
            cp[z]=up[z]/den
            y[z]=(y[z]-lo[z]*y[z-1])/den
        for z in range(L-2,-1,-1):
            y[z]-=cp[z]*y[z+1]
        return y
    nb=mesh.neigh.copy(); has=nb>=0; nbs=np.where(has,nb,0)
    def lat_apply(cols, x):  # sum_j lat_j x[nb_j]
        acc=np.zeros((L,len(cols)))
        for j in range(3):
            acc+=lat[j][:,cols]*np.where(has[cols,j][None,:], x[:,nbs[cols,j]], 0.0)
        return acc
    A=o.suspension_csr(asm); b=rhs.reshape(-1); bn=np.linalg.norm(b)
    def resid(x): return np.linalg.norm(b-A@x.reshape(-1))/bn
    # global multicolour GS
    x=np.zeros((L,T)); kg=None
    for k in range(200):
        for c in (0,1):
            cols=np.where(col==c)[0]
            x[:,cols]=thomas(cols, rhs[:,cols]-lat_apply(cols,x))
        if resid(x)<=1e-8: kg=k+1;break
    print(f"n={n} T={T} global GS sweeps {kg}")
    tid=np.arange(T)//tile
    ntile=tid.max()+1
    for kin in ks:
        x=np.zeros((L,T)); kout=None
        for it in range(200):
            xold=x.copy()
            for kk in range(kin):
                for c in (0,1):
                    cols=np.where(col==c)[0]
                    # neighbours inside own tile use current x, outside use xold
                    acc=np.zeros((L,len(cols)))
                    for j in range(3):
                        n_=nbs[cols,j]; inside=(tid[n_]==tid[cols])
                        xv=np.where(inside[None,:], x[:,n_], xold[:,n_])
                        acc+=lat[j][:,cols]*np.where(has[cols,j][None,:], xv, 0.0)
                    x[:,cols]=thomas(cols, rhs[:,cols]-acc)
            if resid(x)<=1e-8: kout=it+1;break
        print(f"   tile {tile} faces, {kin} inner sweeps: outer iterations {kout}  (coefficient passes {kout}, vs {kg})")
for tile in (256, 512, 1024):
    run(120, tile, (2,4,8,16))
