function [Ainr,binr] = CollConstrHardDMPCOnDemand(p,po,vo,n,k,l,rmin,Ain,A_initp,E1,E2,order,violation)
% Drop-in for dmpc/matlab/CollConstrHardDMPCOnDemand.m:1-32 (rows of the neighbours in `violation`, no slack).
if order ~= 2, error('dmpcb200:order','only order = 2 is implemented'); end
P = struct('N',size(l,3),'K',size(l,2),'variant',3,'rmin',rmin,'c',1/E1(3,3),'h',A_initp(1,4));
[Ainr,binr,~] = dmpc_b200_mex('constr',P,p(:),po(:),vo(:),n,k,l,logical(violation(:)));
end
