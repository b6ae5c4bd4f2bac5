// Reads the input files written by make_inputs.py (canonical little-endian field elements, one file per case),
// commits to each with the unmodified reference and prints one JSON object per case:
//   {"case": "...", "root": "<hex>", "n_rows": .., "n_per_row": .., "n_cols": ..,
//    "proof_len": .., "proof_blake3": "<hex>", "eval": "<hex>"}
// proof_* describe bincode::serialize(&proof) of an evaluation proof made on Transcript::new(b"rust golden") with
// outer tensor = the first n_rows input coefficients; eval = to_repr of what verify() returns with inner tensor =
// the first n_per_row input coefficients.
// Paste the output into tests/golden/rust_roots.json; tests/test_rust_golden.py then holds the oracle to it.
use blake3::Hasher as Blake3;
use ff::PrimeField;
use lcpc_2d::LcEncoding;
use merlin::Transcript;
use lcpc_brakedown_pc::{BrakedownCommit, SdigEncoding};
use lcpc_ligero_pc::{LigeroCommit, LigeroEncoding};
use lcpc_test_fields::{ft127::Ft127, ft255::Ft255, ft63::Ft63};
use std::{env, fs};

fn read_elems<F: PrimeField>(path: &str, bytes_per: usize) -> Vec<F> {
    let raw = fs::read(path).expect("input file");
    raw.chunks(bytes_per)
        .map(|c| {
            let mut repr = F::Repr::default();
            repr.as_mut().copy_from_slice(c); // PrimeFieldReprEndianness = "little"
            Option::<F>::from(F::from_repr(repr)).expect("canonical element")
        })
        .collect()
}

fn hex(b: &[u8]) -> String {
    b.iter().map(|x| format!("{:02x}", x)).collect()
}

macro_rules! ligero_case {
    ($dir:expr, $name:expr, $f:ty, $bytes:expr) => {{
        let coeffs: Vec<$f> = read_elems(&format!("{}/{}.bin", $dir, $name), $bytes);
        let enc = LigeroEncoding::<$f>::new(coeffs.len());
        let (n_rows, n_per_row, n_cols) = enc.get_dims(coeffs.len());
        let comm = LigeroCommit::<Blake3, $f>::commit(&coeffs, &enc).unwrap();
        let root = comm.get_root();
        let outer: Vec<$f> = coeffs[..n_rows].to_vec();
        let inner: Vec<$f> = coeffs[..n_per_row].to_vec();
        let mut tr = Transcript::new(b"rust golden");
        let pf = comm.prove(&outer[..], &enc, &mut tr).unwrap();
        let wire: Vec<u8> = bincode::serialize(&pf).unwrap();
        let mut tr2 = Transcript::new(b"rust golden");
        let ev = pf.verify(root.as_ref(), &outer[..], &inner[..], &enc, &mut tr2).unwrap();
        println!(
            "{{\"case\": \"{}\", \"root\": \"{}\", \"n_rows\": {}, \"n_per_row\": {}, \"n_cols\": {}, \"proof_len\": {}, \"proof_blake3\": \"{}\", \"eval\": \"{}\"}}",
            $name, hex(root.as_ref()), n_rows, n_per_row, n_cols, wire.len(),
            hex(blake3::hash(&wire).as_bytes()), hex(ev.to_repr().as_ref())
        );
    }};
}

macro_rules! brakedown_case {
    ($dir:expr, $name:expr, $f:ty, $bytes:expr, $seed:expr) => {{
        let coeffs: Vec<$f> = read_elems(&format!("{}/{}.bin", $dir, $name), $bytes);
        let enc = SdigEncoding::<$f>::new(coeffs.len(), $seed);
        let (n_rows, n_per_row, n_cols) = enc.get_dims(coeffs.len());
        let comm = BrakedownCommit::<Blake3, $f>::commit(&coeffs, &enc).unwrap();
        let root = comm.get_root();
        let outer: Vec<$f> = coeffs[..n_rows].to_vec();
        let inner: Vec<$f> = coeffs[..n_per_row].to_vec();
        let mut tr = Transcript::new(b"rust golden");
        let pf = comm.prove(&outer[..], &enc, &mut tr).unwrap();
        let wire: Vec<u8> = bincode::serialize(&pf).unwrap();
        let mut tr2 = Transcript::new(b"rust golden");
        let ev = pf.verify(root.as_ref(), &outer[..], &inner[..], &enc, &mut tr2).unwrap();
        println!(
            "{{\"case\": \"{}\", \"root\": \"{}\", \"n_rows\": {}, \"n_per_row\": {}, \"n_cols\": {}, \"proof_len\": {}, \"proof_blake3\": \"{}\", \"eval\": \"{}\"}}",
            $name, hex(root.as_ref()), n_rows, n_per_row, n_cols, wire.len(),
            hex(blake3::hash(&wire).as_bytes()), hex(ev.to_repr().as_ref())
        );
    }};
}

fn main() {
    let dir = env::args().nth(1).expect("usage: lcpc-b200-rust-golden <input dir>");
    ligero_case!(dir, "ligero_ft255_2_10", Ft255, 32);
    ligero_case!(dir, "ligero_ft255_2_14", Ft255, 32);
    ligero_case!(dir, "ligero_ft127_2_12", Ft127, 16);
    ligero_case!(dir, "ligero_ft63_1000", Ft63, 8);
    brakedown_case!(dir, "brakedown_ft127_2_12_seed0", Ft127, 16, 0u64);
    brakedown_case!(dir, "brakedown_ft255_3000_seed1", Ft255, 32, 1u64);
    brakedown_case!(dir, "brakedown_ft63_2_13_seed7", Ft63, 8, 7u64);
}
