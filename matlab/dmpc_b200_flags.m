function [p,v,a,success,outbound,coll] = dmpc_b200_flags(p,v,a,st)
% status word of include/dmpc_b200.h -> the reference's (success/feasible, outbound, coll) flags and
% empty p,v,a where the reference returns [] (solveSoftDMPCbound.m:26-31,136-139).
st = double(st);
solved   = bitand(st,1) ~= 0;
coll     = double(bitand(st,2) ~= 0);
success  = double(bitand(st,4+16) == 0);
outbound = double(bitand(st,8) ~= 0);
if ~solved, p = []; v = []; a = []; end
end
