function inbounds = is_inbounds(p,pmin,pmax)
% Drop-in for dmpc/matlab/is_inbounds.m:1-6.  A three-comparison host test in the reference as well; inside
% the batched step the same test runs on the device (status bit 8 of include/dmpc_b200.h, inb_tol = 50e-3).
tol = 50e-3;
up = max(p(1,:)) < pmax(1)+tol && max(p(2,:)) < pmax(2)+tol && max(p(3,:)) < pmax(3)+tol;
down = min(p(1,:)) > pmin(1)-tol && min(p(2,:)) > pmin(2)-tol && min(p(3,:)) > pmin(3)-tol;
inbounds = up && down;
end
