import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla, sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from chm_b200 import synthetic

def dep_matrix(mesh, eps=6500.0):
    geo=mesh.geometry(); T=mesh.n_local
    rows=[];cols=[];vals=[]
    diag=geo.area.copy()
    for j in range(3):
        n=mesh.neigh[:,j]; has=n>=0
        dx=np.hypot(geo.cx[has]-geo.cx[n[has]], geo.cy[has]-geo.cy[n[has]])
        c=eps*geo.elen[j][has]/dx
        diag[has]+=c
        rows.append(np.where(has)[0]); cols.append(n[has]); vals.append(-c)
    A=sp.csr_matrix((np.concatenate(vals),(np.concatenate(rows),np.concatenate(cols))),shape=(T,T))+sp.diags(diag)
    return A.tocsr(), diag

def colour(neigh):
    T=neigh.shape[0]
    col=-np.ones(T,int)
    # try bipartite BFS
    ok=True
    for s in range(T):
        if col[s]>=0: continue
        col[s]=0; st=[s]
        while st and ok:
            i=st.pop()
            for n in neigh[i]:
                if n<0: continue
                if col[n]<0: col[n]=1-col[i]; st.append(n)
                elif col[n]==col[i]: ok=False;break
        if not ok: break
    if ok: return col,2
    col[:]=-1
    for i in range(T):
        used=set(col[n] for n in neigh[i] if n>=0 and col[n]>=0)
        c=0
        while c in used: c+=1
        col[i]=c
    return col,col.max()+1

def run(mesh,name):
    A,d=dep_matrix(mesh); T=mesh.n_local
    rng=np.random.default_rng(0)
    geo=mesh.geometry()
    b=np.sin(geo.cx/700.0)*np.cos(geo.cy/500.)*geo.area + 0.1*rng.standard_normal(T)
    Dinv=1.0/d
    J=sp.eye(T)-sp.diags(Dinv)@A
    lmax=spla.eigsh(sp.diags(Dinv**0.5)@A@sp.diags(Dinv**0.5),k=1,which='LA',return_eigenvectors=False)[0]
    lmin=spla.eigsh(sp.diags(Dinv**0.5)@A@sp.diags(Dinv**0.5),k=1,sigma=0,which='LM',return_eigenvectors=False)[0]
    kappa=lmax/lmin
    rho=1-lmin
    # chebyshev jacobi
    theta=(lmax+lmin)/2; delta=(lmax-lmin)/2; sigma1=theta/delta
    q=np.zeros(T); dvec=np.zeros(T); rhoc=1/sigma1; bn=np.linalg.norm(b)
    kc=None
    for k in range(2000):
        r=b-A@q
        if np.linalg.norm(r)<=1e-8*bn: kc=k;break
        z=Dinv*r
        if k==0: dvec=z/theta
        else:
            rn=1/(2*sigma1-rhoc); dvec=rn*rhoc*dvec+2*rn/delta*z; rhoc=rn
        q=q+dvec
    col,nc=colour(mesh.neigh)
    masks=[np.where(col==c)[0] for c in range(nc)]
    Ac=[A[m] for m in masks]
    res={}
    for w in [1.0, 2/(1+np.sqrt(1-rho**2)), 2/(1+np.sqrt(1-rho**2))*0.98, 2/(1+np.sqrt(1-rho**2))*1.01]:
        q=np.zeros(T); ks=None
        for k in range(2000):
            for m,Am in zip(masks,Ac):
                r=b[m]-Am@q
                q[m]+=w*Dinv[m]*r
            if np.linalg.norm(b-A@q)<=1e-8*bn: ks=k+1;break
        res[round(w,4)]=ks
    print(name,"T",T,"colours",nc,"kappa %.0f"%kappa,"cheb",kc,"SOR",res)

run(synthetic.uniform_mesh(100,100),"uniform100")
run(synthetic.uniform_mesh(200,200),"uniform200")
run(synthetic.variable_mesh(20000),"var20k")
run(synthetic.variable_mesh(80000),"var80k")

def run_part(mesh,name,P,wscale=1.0,mode="sweep"):
    """rank-local colouring, ghosts refreshed once per sweep (mode sweep) or after each local colour index pass (mode colour)"""
    A,d=dep_matrix(mesh); T=mesh.n_local
    rng=np.random.default_rng(0); geo=mesh.geometry()
    b=np.sin(geo.cx/700.0)*np.cos(geo.cy/500.)*geo.area + 0.1*rng.standard_normal(T)
    Dinv=1.0/d; bn=np.linalg.norm(b)
    S=sp.diags(Dinv**0.5)@A@sp.diags(Dinv**0.5)
    lmin=spla.eigsh(S,k=1,sigma=0,which='LM',return_eigenvectors=False)[0]
    rho=1-lmin; w=2/(1+np.sqrt(1-rho**2))*wscale
    bounds=[T*r//P for r in range(P+1)]
    colg=-np.ones(T,int); ncmax=0
    for r in range(P):
        s,e=bounds[r],bounds[r+1]
        nl=mesh.neigh[s:e].copy(); nl=np.where((nl>=s)&(nl<e), nl-s, -1)
        c,nc=colour(nl); colg[s:e]=c; ncmax=max(ncmax,nc)
    rank=np.searchsorted(bounds,np.arange(T),side='right')-1
    q=np.zeros(T); ks=None
    A=A.tocsr()
    for k in range(3000):
        qold=q.copy()   # ghost values as of sweep start
        for c in range(ncmax):
            m=np.where(colg==c)[0]
            if mode=="colour": qold=q.copy()
            # own-rank neighbours use current q, other-rank neighbours use qold: build mixed vector per rank
            newvals=np.empty(len(m))
            for r in range(P):
                mr=m[rank[m]==r]
                if len(mr)==0: continue
                qm=qold.copy(); s,e=bounds[r],bounds[r+1]; qm[s:e]=q[s:e]
                rr=b[mr]-A[mr]@qm
                q[mr]=q[mr]+w*Dinv[mr]*rr
        if not np.isfinite(q).all(): ks="diverged";break
        if np.linalg.norm(b-A@q)<=1e-8*bn: ks=k+1;break
    print(name,"P",P,"mode",mode,"w %.4f"%w,"colours",ncmax,"SOR sweeps",ks)

m=synthetic.uniform_mesh(100,100)
for P in (1,2,8):
    run_part(m,"uniform100",P)
run_part(m,"uniform100",8,mode="colour")
m=synthetic.variable_mesh(20000)
for P in (1,2,8):
    run_part(m,"var20k",P)
run_part(m,"var20k",8,mode="colour")
run_part(m,"var20k",8,wscale=1.01)
