function [Ain_total,bin_total] = CollConstrHardDMPC(p,po,vo,n,k,l,rmin,Ain,A_initp,E1,E2,order)
% Drop-in for dmpc/matlab/CollConstrHardDMPC.m:1-34.  The reference returns N-1 rows: one per neighbour with
% dist < 1 (:19), packed first in ascending neighbour order, and ALL-ZERO rows (0*x <= 0) for the rest (:3-5).
% The device returns the nr non-trivial rows; they are padded here so that size(Ain_total) = [N-1, 3K] as in
% the reference (solveHardDMPC.m:18-22 stacks K of these blocks).
if order ~= 2, error('dmpcb200:order','only order = 2 is implemented'); end
N = size(l,3); K = size(l,2);
P = struct('N',N,'K',K,'variant',2,'rmin',rmin,'c',1/E1(3,3),'h',A_initp(1,4));
[A,b,~] = dmpc_b200_mex('constr',P,p(:),po(:),vo(:),n,k,l,[]);
Ain_total = zeros(N-1,3*K); bin_total = zeros(N-1,1);
nr = size(A,1);
Ain_total(1:nr,:) = A; bin_total(1:nr) = b;
end
