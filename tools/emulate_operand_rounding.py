"""CPU emulation (float64 + explicit 16-bit roundings) of the two ways the backward can treat the target term
"- 2 T_ij": old = round G~ to 16 bit, subtract 2 Q in fp32; new = round G~ - lam2 (loss_bwd_pair.cu epilogue),
subtract (2 - lam2) Q in fp32.  Prints the relative gradient error of both against exact arithmetic: the floor
the GPU tests of the trained regime (tests/test_loss_gpu.py) are held to.  Development aid, not product code."""
import torch, numpy as np
torch.manual_seed(0)
def rnd(x, kind):
    return x.bfloat16().double() if kind=="bf16" else x.half().double()
def emu(feats, labels, s, kind, scheme):
    # feats: list of 2 double tensors [N,d] (one pair a,b); returns dXhat_a (gradient wrt unit rows), emulating roundings
    A,B = feats
    N = A.shape[0]
    Ah = A/A.norm(dim=1,keepdim=True); Bh = B/B.norm(dim=1,keepdim=True)
    Ar, Br = rnd(Ah.float(),kind), rnd(Bh.float(),kind)
    S = Ar@Br.T
    E = torch.exp(s*S - s)
    T = (labels[:,None]==labels[None,:]).double()
    c = T.sum(1)
    r = E.sum(1); q = E.sum(0)
    u = c/r; v = c/q
    G = E*(u[:,None]+v[None,:])
    if scheme=="old":
        Gr = rnd(G.float(),kind)
        dA = Gr@Br - 2*(T@Bh)
    else:
        pos = (Ah*(T@Bh)).sum(1)
        lam2 = 2*torch.clamp(0.5*torch.exp(s*(pos/c-1))*(u+v), max=1.0)
        Gr = rnd((G - lam2[:,None]*T).float(),kind)
        dA = Gr@Br - (2-lam2)[:,None]*(T@Bh)
    # exact
    Sx = Ah@Bh.T
    # exact computed on same (unrounded) inputs
    Ex = torch.exp(s*Sx-s); rx=Ex.sum(1); qx=Ex.sum(0)
    Gx = Ex*((c/rx)[:,None]+(c/qx)[None,:]) - 2*T
    dAx = Gx@Bh
    return ((dA-dAx).norm()/dAx.norm()).item()
if __name__ == "__main__":
    N,d=768,768
    for labels_kind, align in [("onehot",0.7),("multi",0.7),("multi",0.0),("onehot",0.0)]:
        gen=torch.Generator().manual_seed(21)
        labels = torch.arange(N) if labels_kind=="onehot" else torch.randint(0,N//4,(N,),generator=gen)
        centres=torch.randn(N,d,generator=gen)
        base = centres[labels] if labels_kind=="multi" else centres
        for kind in ("bf16","fp16"):
            feats=[rnd(align*base+(1-align)*torch.randn(N,d,generator=gen),kind) for _ in range(2)]
            print(labels_kind, align, kind, "old", emu(feats,labels,1/0.07,kind,"old"), "new", emu(feats,labels,1/0.07,kind,"new"))
