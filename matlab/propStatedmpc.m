function [p,v] = propStatedmpc(po,vo,a,A_initp,A_p,A_v)
% Drop-in for dmpc/matlab/propStatedmpc.m.
K = length(a)/3;
P = struct('N',1,'K',K,'h',A_initp(1,4));
[p,v] = dmpc_b200_mex('prop',P,po(:),vo(:),a(:));
end
