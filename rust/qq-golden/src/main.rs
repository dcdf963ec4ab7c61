//! Golden vectors from the reference (see Cargo.toml).  NOT BUILT HERE.
//!
//! File = records `tag: u32 LE | len: u64 LE | payload`:
//!   1  update_account      n x (Account 128 | bl 32 | u 32 | c 32 | Account::update_account(..) 128)
//!   2  delta / epsilon     n: u64 | accounts n x 128 | bl n x 32 | r n x 32 | delta n x 128 | epsilon n x 128
//!   3  shuffle proof       four length-prefixed (u64) bincode blobs: Vec<Account> inputs, Vec<Account> outputs,
//!                          ShuffleStatement, ShuffleProof      (transcript b"ShuffleProof", verifier b"Shuffle")
//!   4  DLOG sigma proof    three length-prefixed bincode blobs: Vec<Account> updated, Vec<Account> updated_delta,
//!                          SigmaProof                          (transcript b"UpdateAccount", verifier b"DLOGProof")
//!   5  range proof         m: u32 | commitments m x 32 | RangeProof::to_bytes()   (Transcript::new(b"qq-golden"), n = 64)
//!   6  BulletproofGens     BulletproofGens::new(64, 16): G[0..4) then H[0..4) of party 0 and of party 1, 32 B each
//!   7  VectorPedersenGens  VectorPedersenGens::new(4): H | G[0..3)
#![allow(non_snake_case)]
use bulletproofs::{BulletproofGens, PedersenGens, RangeProof};
use curve25519_dalek::scalar::Scalar;
use merlin::Transcript;
use quisquislib::accounts::{Account, Prover, Verifier};
use quisquislib::keys::{PublicKey, SecretKey};
use quisquislib::pedersen::vectorpedersen::VectorPedersenGens;
use quisquislib::ristretto::{RistrettoPublicKey, RistrettoSecretKey};
use quisquislib::shuffle::{Shuffle, ShuffleProof};
use std::io::Write;

fn record(out: &mut Vec<u8>, tag: u32, payload: &[u8]) {
    out.extend_from_slice(&tag.to_le_bytes());
    out.extend_from_slice(&(payload.len() as u64).to_le_bytes());
    out.extend_from_slice(payload);
}
fn blob(out: &mut Vec<u8>, b: &[u8]) {
    out.extend_from_slice(&(b.len() as u64).to_le_bytes());
    out.extend_from_slice(b);
}
fn acc_bytes(a: &Account) -> Vec<u8> {
    bincode::serialize(a).unwrap() // pk.gr | pk.grsk | comm.c | comm.d, 128 bytes
}
fn random_account(value: u64) -> (Account, RistrettoSecretKey) {
    Account::generate_random_account_with_value(Scalar::from(value))
}

fn main() {
    let path = std::env::args().nth(1).expect("output file");
    let mut rng = rand::thread_rng();
    let mut out = Vec::new();

    // 1: update_account (src/accounts/accounts.rs:143-154)
    let mut p = Vec::new();
    for i in 0..64u64 {
        let (acc, _) = random_account(i % 7);
        let bl = if i % 3 == 0 { Scalar::from(i) } else if i % 3 == 1 { -Scalar::from(i) } else { Scalar::random(&mut rng) };
        let (u, c) = (Scalar::random(&mut rng), Scalar::random(&mut rng));
        let upd = Account::update_account(acc, bl, u, c);
        p.extend(acc_bytes(&acc));
        p.extend_from_slice(bl.as_bytes());
        p.extend_from_slice(u.as_bytes());
        p.extend_from_slice(c.as_bytes());
        p.extend(acc_bytes(&upd));
    }
    record(&mut out, 1, &p);

    // 2: create_delta_and_epsilon_accounts (src/accounts/accounts.rs:198-220); r comes back from the reference
    let accounts: Vec<Account> = (0..9).map(|i| random_account(i).0).collect();
    let bl: Vec<Scalar> = vec![-Scalar::from(5u64), Scalar::from(5u64)].into_iter().chain((0..7).map(|_| Scalar::zero())).collect();
    let base_pk = RistrettoPublicKey::generate_base_pk();
    let (delta, epsilon, r) = Account::create_delta_and_epsilon_accounts(&accounts, &bl, base_pk);
    let mut p = Vec::new();
    p.extend_from_slice(&(accounts.len() as u64).to_le_bytes());
    accounts.iter().for_each(|a| p.extend(acc_bytes(a)));
    bl.iter().for_each(|s| p.extend_from_slice(s.as_bytes()));
    r.iter().for_each(|s| p.extend_from_slice(s.as_bytes()));
    delta.iter().for_each(|a| p.extend(acc_bytes(a)));
    epsilon.iter().for_each(|a| p.extend(acc_bytes(a)));
    record(&mut out, 2, &p);

    // 3: a shuffle proof, serialised as the reference serialises it (src/shuffle/shuffle.rs:759-795)
    for _ in 0..2 {
        let mut account_vector: Vec<Account> = Vec::new();
        for _ in 0..9 {
            let sk: RistrettoSecretKey = SecretKey::random(&mut rng);
            let pk = RistrettoPublicKey::from_secret_key(&sk, &mut rng);
            account_vector.push(Account::generate_account(pk).0);
        }
        let shuffle = Shuffle::input_shuffle(&account_vector).unwrap();
        let pc_gens = PedersenGens::default();
        let xpc_gens = VectorPedersenGens::new(4);
        let mut transcript_p = Transcript::new(b"ShuffleProof");
        let mut prover = Prover::new(b"Shuffle", &mut transcript_p);
        let (proof, statement) = ShuffleProof::create_shuffle_proof(&mut prover, &shuffle, &pc_gens, &xpc_gens);
        let mut transcript_v = Transcript::new(b"ShuffleProof");
        let mut verifier = Verifier::new(b"Shuffle", &mut transcript_v);
        assert!(proof
            .verify(&mut verifier, &statement, &shuffle.get_inputs_vector(), &shuffle.get_outputs_vector(), &pc_gens, &xpc_gens)
            .is_ok());
        let mut p = Vec::new();
        blob(&mut p, &bincode::serialize(&shuffle.get_inputs_vector()).unwrap());
        blob(&mut p, &bincode::serialize(&shuffle.get_outputs_vector()).unwrap());
        blob(&mut p, &bincode::serialize(&statement).unwrap());
        blob(&mut p, &bincode::serialize(&proof).unwrap());
        record(&mut out, 3, &p);
    }

    // 4: DLOG proof of a correct account update (src/accounts/verifier.rs:1006-1072)
    {
        let updated: Vec<Account> = (0..9).map(|_| random_account(0).0).collect();
        let values: Vec<Scalar> = vec![Scalar::zero(); 9];
        let (delta, _, rscalars) = Account::create_delta_and_epsilon_accounts(&updated, &values, RistrettoPublicKey::generate_base_pk());
        let updated_delta = Account::update_delta_accounts(&updated, &delta).unwrap();
        let mut transcript = Transcript::new(b"UpdateAccount");
        let mut prover = Prover::new(b"DLOGProof", &mut transcript);
        let sigma = Prover::verify_update_account_prover(&updated[2..9], &updated_delta[2..9], &rscalars[2..9], &mut prover);
        let mut p = Vec::new();
        blob(&mut p, &bincode::serialize(&updated[2..9].to_vec()).unwrap());
        blob(&mut p, &bincode::serialize(&updated_delta[2..9].to_vec()).unwrap());
        blob(&mut p, &bincode::serialize(&sigma).unwrap());
        record(&mut out, 4, &p);
    }

    // 5: aggregated range proofs from the bulletproofs crate itself, m = 1, 4, 16 (src/accounts/verifier.rs:504-523 calls
    // verify_multiple with these generators)
    let pc_gens = PedersenGens::default();
    let bp_gens = BulletproofGens::new(64, 16);
    for &m in &[1usize, 4, 16] {
        let values: Vec<u64> = (0..m as u64).map(|i| 1_000_000 * (i + 1) + 17).collect();
        let blindings: Vec<Scalar> = (0..m).map(|_| Scalar::random(&mut rng)).collect();
        let mut t = Transcript::new(b"qq-golden");
        let (proof, commitments) = RangeProof::prove_multiple(&bp_gens, &pc_gens, &mut t, &values, &blindings, 64).unwrap();
        let mut tv = Transcript::new(b"qq-golden");
        assert!(proof.verify_multiple(&bp_gens, &pc_gens, &mut tv, &commitments, 64).is_ok());
        let mut p = Vec::new();
        p.extend_from_slice(&(m as u32).to_le_bytes());
        commitments.iter().for_each(|c| p.extend_from_slice(c.as_bytes()));
        p.extend(proof.to_bytes());
        record(&mut out, 5, &p);
    }

    // 6, 7: generator chains
    let mut p = Vec::new();
    for party in 0..2 {
        let share = bp_gens.share(party);
        share.G(4).for_each(|g| p.extend_from_slice(g.compress().as_bytes()));
        share.H(4).for_each(|h| p.extend_from_slice(h.compress().as_bytes()));
    }
    record(&mut out, 6, &p);
    let xpc = VectorPedersenGens::new(4);
    let mut p = Vec::new();
    p.extend_from_slice(xpc.H.compress().as_bytes());
    xpc.G_vec.iter().take(3).for_each(|g| p.extend_from_slice(g.compress().as_bytes()));
    record(&mut out, 7, &p);

    std::fs::File::create(&path).unwrap().write_all(&out).unwrap();
    println!("wrote {} bytes to {}", out.len(), path);
}
