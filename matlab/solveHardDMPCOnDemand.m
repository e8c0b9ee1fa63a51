function [p,v,a,success,outbound,coll] = solveHardDMPCOnDemand(po,pf,vo,ao,n,h,l,K,rmin,pmin,pmax,alim,A,A_initp,A_p,A_v,Delta,Q1,S1,E1,E2,order)
% Drop-in for dmpc/matlab/solveHardDMPCOnDemand.m.
if order ~= 2, error('dmpcb200:order','only order = 2 is implemented'); end
P = struct('N',size(l,3),'K',K,'variant',3,'h',h,'rmin',rmin,'c',1/E1(3,3),'alim',alim,'Q1',Q1,'S1',S1);
[p,v,a,st] = dmpc_b200_mex('solve',P,po(:),pf(:),vo(:),ao(:),n,l,pmin(:),pmax(:));
[p,v,a,success,outbound,coll] = dmpc_b200_flags(p,v,a,st,P.variant);
end
