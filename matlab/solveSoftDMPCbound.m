function [p,v,a,success,outbound,coll] = solveSoftDMPCbound(po,pf,vo,ao,n,h,l,K,rmin,pmin,pmax,alim,A,A_initp,A_p,A_v,Delta,Q1,S1,E1,E2,order,term)
% Drop-in for dmpc/matlab/solveSoftDMPCbound.m (same 23 inputs / 6 outputs) running on B200 through
% dmpc_b200_mex.  A, A_initp, A_p, A_v, Delta, E2 are accepted for signature parity and unused:
% the device evaluates the same constant maps from h and K.  For throughput call dmpc_step.m once
% per MPC step instead of this function once per agent.
if order ~= 2, error('dmpcb200:order','only order = 2 is implemented'); end
P = struct('N',size(l,3),'K',K,'variant',0,'h',h,'rmin',rmin,'c',1/E1(3,3),'alim',alim,'Q1',Q1,'S1',S1,'term',term);
[p,v,a,st] = dmpc_b200_mex('solve',P,po(:),pf(:),vo(:),ao(:),n,l,pmin(:),pmax(:));
[p,v,a,success,outbound,coll] = dmpc_b200_flags(p,v,a,st);
end
