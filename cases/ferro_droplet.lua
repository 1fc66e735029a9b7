-- Ferrofluid droplet in a uniform far field: the magnetostatics Poisson case.
-- Same keys and values as the reference's examples/magnetostatics/magnetostatics.lua, so the
-- reference's own file can be passed instead of this one, unchanged.

-- geometry
dim     = 2
length  = RealVec2(301.0, 301.0)
n_cells = UIntVec2(301, 301)

-- model parameters (examples/magnetostatics/model.rs: UserModel)
model = Model({
    H_far         = RealVec2(0.0, 1.0),        -- far-field H
    bubble_centre = RealVec2(301.0/2, 301.0/2),
    bubble_radius = 301.0/20,
    mu            = RealVec2(3.2, 1.0),        -- inside, outside
    relax         = 0.2,
    n_iter        = 10,                        -- outer under-relaxed iterations
    n_sub_iter    = 550,                       -- relaxation sweeps per outer iteration
    tol           = 0.1
})
