function pass = ReachedGoal(p,pf,length_t,error_tol,N)
% Drop-in for dmpc/matlab/ReachedGoal.m:1-11: max_n ||p(:,length_t,n) - pf(:,n)|| < error_tol, reduced on the
% device (dmpcb200_reached_goal).  Inside dmpc_b200_mex('run', ...) the same test runs after every step
% without leaving the GPU.
if N > 1, pk = squeeze(p(:,length_t,:)); else, pk = p(:,length_t); end
[pass,~] = dmpc_b200_mex('goal',struct('N',N),pk,reshape(pf,3,N),error_tol);
pass = logical(pass);
end
