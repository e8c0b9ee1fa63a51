function [Ainr,binr,prev_dist] = CollConstrSoftDMPC2(p,po,vo,n,k,l,rmin,Ain,A_initp,E1,E2,order,violation)
% Drop-in for dmpc/matlab/CollConstrSoftDMPC2.m:1-32 (k_ctr = k-1, :8).
if order ~= 2, error('dmpcb200:order','only order = 2 is implemented'); end
P = struct('N',size(l,3),'K',size(l,2),'variant',1,'rmin',rmin,'c',1/E1(3,3),'h',A_initp(1,4));
[Ainr,binr,prev_dist] = dmpc_b200_mex('constr',P,p(:),po(:),vo(:),n,k,l,logical(violation(:)));
end
