function [Ain_total,bin_total] = CollConstr(p,po,k,l,Ain,rmin,E1,E2,order)
% The dec-iSCP helper NAME (dec-iSCP/CollConstr.m:1-23) kept callable on the device path.  Its rows are the
% DMPC rows of CollConstrSoftDMPC2 with vo = 0 and no own agent: one row per column of l (all of them are
% obstacles), acting on block k-1 (:17), r = dist*(rmin - dist + diff*p/dist) - diff*po' (:14).  Ain must be
% the position matrix getPosMat(h,K) (as in dec-iSCP/singleiSCP.m:9); h is read from it (Ain(1,1) = h^2/2).
if order ~= 2, error('dmpcb200:order','only order = 2 is implemented'); end
if k < 2, error('dmpcb200:arg','CollConstr needs k >= 2 (the row acts on block k-1)'); end
N = size(l,3); K = size(l,2);
P = struct('N',N,'K',K,'variant',1,'rmin',rmin,'c',1/E1(3,3),'h',sqrt(2*Ain(1,1)));
[Ain_total,bin_total,~] = dmpc_b200_mex('constr',P,p(:),po(:),zeros(3,1),0,k,l,true(N,1),N);
end
