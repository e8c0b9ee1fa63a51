function [p,v] = propState(po,a,A_p,A_v,K)
% The dec-iSCP helper NAME (dec-iSCP/propState.m:1-10) kept callable on the device path: K trajectory points
% from rest, p = [po; A_p*a + po], v = [0; A_v*a] with A_p, A_v the first 3(K-1) rows of the kinematic maps
% (dec-iSCP/decSCP.m:55-71) -- i.e. the first K-1 rows of propStatedmpc.m with vo = 0.
h = A_v(1,1);
[pp,vv] = dmpc_b200_mex('prop',struct('N',1,'K',K,'h',h),po(:),zeros(3,1),a(:));
p = [po(:); pp(1:3*(K-1))];
v = [zeros(3,1); vv(1:3*(K-1))];
end
