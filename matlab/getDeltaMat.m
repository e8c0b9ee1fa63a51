function Delta = getDeltaMat(K)
% Drop-in for dmpc/matlab/getDeltaMat.m.
[~,~,~,Delta] = dmpc_b200_mex('mats',1,K);
end
