function [p,v,a,success,outbound,coll] = dmpc_b200_flags(p,v,a,st,variant)
% status word of include/dmpc_b200.h -> the reference's (success/feasible, outbound, coll) flags and
% empty p,v,a where the reference returns [] (solveSoftDMPCbound.m:26-31,136-139).
% variant 0 (solveSoftDMPCbound.m) keeps feasible = 1 on the k == 1 collision exit (:29-30) and when the
% first position is out of bounds (:125-128); variants 1..3 (solveSoftDMPCbound2.m:14,26,123-126,
% solveHardDMPC.m:14,76-79, solveHardDMPCOnDemand.m:14,81-84) return success = 0 in both cases.
% Internal failures of the library (16 iteration cap, 32 capacity overflow) are "not feasible" everywhere.
if nargin < 5, variant = 0; end
st = double(st);
solved   = bitand(st,1) ~= 0;
coll     = double(bitand(st,2) ~= 0);
outbound = double(bitand(st,8) ~= 0);
hardfail = bitand(st,4+16+32) ~= 0;
if variant == 0
    success = double(~hardfail);
else
    success = double(solved && ~outbound && ~coll && ~hardfail);
end
if ~solved, p = []; v = []; a = []; end
end
