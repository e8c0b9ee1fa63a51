function A = getPosMat(h,K)
% Drop-in for dmpc/matlab/getPosMat.m (bit-identical to the reference's recurrence).
[A,~,~,~] = dmpc_b200_mex('mats',h,K);
end
