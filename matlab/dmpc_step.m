function [new_l,pk1,vk1,ak1,status,first_fail] = dmpc_step(P,pk,vk,ak,l)
% ONE batched MPC step for all N agents: replaces the whole `for n = 1:N` body of
% test/failure_rate.m:100-119 / dmpc_soft_bound.m:116-135 (Jacobi: every agent reads l of the
% previous step).  pk,vk,ak are 3 x N (column k-1 of the reference's pk,vk,ak), l is 3 x K x N.
% Agents that fail keep their horizon and state; first_fail is the lowest failing agent (0: none).
[new_l,pk1,vk1,ak1,status,first_fail] = dmpc_b200_mex('step',P,pk,vk,ak,l);
end
