//! CPU baseline with the reference's own code (see Cargo.toml).  NOT BUILT HERE.
use curve25519_dalek::ristretto::{CompressedRistretto, RistrettoPoint};
use curve25519_dalek::scalar::Scalar;
use curve25519_dalek::traits::VartimeMultiscalarMul;
use quisquislib::accounts::Account;
use rayon::prelude::*;
use std::time::Instant;

fn main() {
    let args: Vec<usize> = std::env::args().skip(1).map(|a| a.parse().unwrap()).collect();
    let n_accounts = *args.get(0).unwrap_or(&(1 << 16));
    let n_points = *args.get(1).unwrap_or(&(1 << 20));
    let cores = rayon::current_num_threads();
    let mut rng = rand::thread_rng();

    // BASELINE configs[1]: Account::update_account over n accounts, rayon::par_iter over all host cores
    let accounts: Vec<Account> = (0..n_accounts).map(|i| Account::generate_random_account_with_value(Scalar::from((i % 7) as u64)).0).collect();
    let scalars: Vec<(Scalar, Scalar, Scalar)> =
        (0..n_accounts).map(|_| (Scalar::random(&mut rng), Scalar::random(&mut rng), Scalar::random(&mut rng))).collect();
    let t = Instant::now();
    let updated: Vec<Account> =
        accounts.par_iter().zip(scalars.par_iter()).map(|(a, (bl, u, c))| Account::update_account(*a, *bl, *u, *c)).collect();
    let dt = t.elapsed().as_secs_f64();
    println!(
        "{{\"impl\": \"reference\", \"kind\": \"reference\", \"metric\": \"account_updates_per_sec\", \"value\": {:.1}, \"cores\": {}, \"accounts\": {}}}",
        n_accounts as f64 / dt, cores, updated.len()
    );

    // BASELINE configs[3] / [4]: one vartime MSM over compressed points (decompression included, as the reference's
    // Verifier::multiscalar_multiplication does), the terms split over the cores and the partial sums added
    let points: Vec<CompressedRistretto> = (0..n_points).map(|_| RistrettoPoint::random(&mut rng).compress()).collect();
    let ss: Vec<Scalar> = (0..n_points).map(|_| Scalar::random(&mut rng)).collect();
    let chunk = (n_points + cores - 1) / cores;
    let t = Instant::now();
    let total: RistrettoPoint = points
        .par_chunks(chunk)
        .zip(ss.par_chunks(chunk))
        .map(|(p, s)| RistrettoPoint::optional_multiscalar_mul(s.iter(), p.iter().map(|c| c.decompress())).unwrap())
        .reduce(|| RistrettoPoint::default(), |a, b| a + b);
    let dt = t.elapsed().as_secs_f64();
    println!(
        "{{\"impl\": \"reference\", \"kind\": \"reference\", \"metric\": \"msm_points_per_sec\", \"value\": {:.1}, \"cores\": {}, \"points\": {}, \"check\": \"{:02x}\"}}",
        n_points as f64 / dt, cores, n_points, total.compress().as_bytes()[0]
    );
}
