//! The outer loop of examples/magnetostatics/main.rs:63-74 on device frames: what the re-bodied
//! `get_source / solve_poisson / sum` and the frame operators do underneath.  (`cargo run --example
//! magnetostatics_loop` with RNMF_B200_LIB_DIR set; values of magnetostatics.lua.)
use rnmf_b200::{gs_sweeps, poisson_coefs, poisson_source_2d, BCs, BcType, ComponentBCs, Context, DeviceFrame, MuParams};

fn main() {
    let (n, length) = ([301i64, 301], [301.0f64, 301.0]);
    let dx = [length[0] / n[0] as f64, length[1] / n[1] as f64];
    // examples/magnetostatics/magnetostatics.lua:7-16: H_far (0, 1), bubble_centre 301/2, bubble_radius 301/20, mu (3.2, 1.0),
    // relax 0.2, n_iter 10, n_sub_iter 550, tol 0.1
    let (h_far, relax, n_iter, n_sub_iter, tol) = ([0.0, 1.0], 0.2, 10usize, 550usize, 0.1);
    let mu = MuParams { bubble_centre: [301.0 / 2.0, 301.0 / 2.0], bubble_radius: 301.0 / 20.0, mu: [3.2, 1.0] };
    let (gx, gy) = (-h_far[0] / dx[0], -h_far[1] / dx[1]);
    let bc = BCs { bcs: vec![ComponentBCs { lo: vec![BcType::Neumann(gx), BcType::Neumann(gy)], hi: vec![BcType::Neumann(gx), BcType::Neumann(gy)] }] };

    let ctx = Context::new(0);
    let mut psi = DeviceFrame::new(&ctx, &n, &[0.0, 0.0], &dx, &bc, 1, 1);
    // fill_ic (main.rs:50-53) at cell centres lo + (i+0.5)dx, flat order i fastest
    let mut ic = Vec::with_capacity((n[0] * n[1]) as usize);
    for j in 0..n[1] {
        for i in 0..n[0] {
            let (x, y) = ((i as f64 + 0.5) * dx[0], (j as f64 + 0.5) * dx[1]);
            ic.push(-h_far[0] / dx[0] * x - h_far[1] / dx[1] * y);
        }
    }
    psi.upload_valid(&ic);
    psi.fill_bc();
    let (mut rhs, mut psi_star, mut diff) = (psi.like(&bc), psi.like(&bc), psi.like(&bc));
    let coef = poisson_coefs(&dx);
    let (mut iter, mut sum_diff) = (0, 1.0);
    while iter < n_iter && sum_diff > tol {
        poisson_source_2d(&psi, &mu, relax, &mut rhs); // get_source
        psi_star.copy_from(&psi); // solve_poisson: clone + sweeps in the reference's order
        gs_sweeps(&mut psi_star, &rhs, &coef, n_sub_iter);
        diff.assign_axpby(1.0, &psi_star, -1.0, &psi); // &psi_star + &(-1.0 * &psi)
        sum_diff = ctx.allreduce_sum(diff.abs_sum()); // sum(&diff)
        let mut next = psi.like(&bc);
        next.assign_axpby(1.0, &psi, relax, &diff); // &psi + &(relax * &diff)
        psi = next;
        psi.fill_bc();
        iter += 1;
        println!("iter {:3}  sum_diff = {:.17e}", iter, sum_diff);
    }
}
