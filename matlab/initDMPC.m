function [p,v,a] = initDMPC(po,pf,h,k_hor,K)
% Drop-in for dmpc/matlab/initDMPC.m:1-13 (straight-line first horizon p = po + t*(pf-po)/10, v = a = 0;
% the fifth argument is unused by the reference too).  One agent per call like the reference; for all
% agents at once use dmpc_b200_mex('init', ...).
P = struct('N',1,'K',k_hor,'h',h);
[l,~,~,~] = dmpc_b200_mex('init',P,po(:),pf(:),[-1e9;-1e9;-1e9],[1e9;1e9;1e9]);
p = l(:,:,1); v = zeros(3,k_hor); a = zeros(3,k_hor);
end
