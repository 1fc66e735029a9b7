-- A small droplet case for quick runs (33 x 33 cells); exercises comments, arithmetic,
-- variables and both call syntaxes of the Lua subset reader.
--[[ block comment:
     ignored ]]
local N = 33
dim = 2
length = RealVec2(N * 1.0, N)
n_cells = UIntVec2(N, 30 + 3)
half = N / 2
model = Model{
    H_far = RealVec2(0.0, 1.0),
    bubble_centre = RealVec2(half, half),
    bubble_radius = 2 ^ 2,
    mu = RealVec2(3.2, 1.0);
    n_iter = 4, n_sub_iter = 5 * 5,
    relax = 1 / 5,
    tol = 1e-1,
}
