function [violation,min_dist,viol_constr] = CheckCollSoftDMPC(p,l,n,k,E1,rmin,order)
% Drop-in for dmpc/matlab/CheckCollSoftDMPC.m.
if order ~= 2, error('dmpcb200:order','only order = 2 is implemented'); end
P = struct('N',size(l,3),'K',size(l,2),'rmin',rmin,'c',1/E1(3,3));
[violation,min_dist,viol_constr] = dmpc_b200_mex('check',P,p(:),l,n,k);
end
