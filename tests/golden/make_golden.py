#!/usr/bin/env python3
"""Collect the reference's own golden vectors for the plonky2 path into tests/golden/ (data only).

Run in the build container (needs /root/reference).  Nothing at test time reads /root/reference.
Sources:
  G5  gnark-plonky2-verifier/testdata/test_circuit/{common_circuit_data,verifier_only_circuit_data,proof_with_public_inputs}.json
  G6  near_bft_finality/proofs/{random/CGZP..,epoch/4RjX..,epoch/CbAH..}/{proof.bin,common_data.json,verifier_data.json,hash.json}
      (proof.json is NOT copied: tests/test_oracle_proof_io.py checked bin<->json equality here and records the sha256)
  KATs (numbers quoted from the reference tests) -> kats.json
"""
import gzip, json, os, shutil, hashlib, sys
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

def gz_copy(src, dst):
    with open(src, "rb") as f, gzip.GzipFile(dst, "wb", mtime=0) as g:
        g.write(f.read())

os.makedirs(f"{OUT}/gnark_test_circuit", exist_ok=True)
for n in ("common_circuit_data.json", "verifier_only_circuit_data.json"):
    shutil.copy(f"{REF}/gnark-plonky2-verifier/testdata/test_circuit/{n}", f"{OUT}/gnark_test_circuit/{n}")
gz_copy(f"{REF}/gnark-plonky2-verifier/testdata/test_circuit/proof_with_public_inputs.json",
        f"{OUT}/gnark_test_circuit/proof_with_public_inputs.json.gz")

json_sha = {}
for kind, name in (("random", "CGZPhFRkL3NvmGaXWBc6N7qJD519EUe6vyNpaEyDe2Ev"),
                   ("epoch", "4RjXBrNcu39wutFTuFpnRHgNqgHxLMcGBKNEQdtkSBhy"),
                   ("epoch", "CbAHBGJ8VQot2m6KhH9PLasMgcDtkPJBfp9bjAEMJ8UK")):
    src = f"{REF}/near_bft_finality/proofs/{kind}/{name}"
    dst = f"{OUT}/near_final_proofs/{kind}_{name[:8]}"
    os.makedirs(dst, exist_ok=True)
    for n in ("proof.bin", "common_data.json", "verifier_data.json", "hash.json"):
        shutil.copy(f"{src}/{n}", f"{dst}/{n}")
    # canonical re-serialisation of proof.json (what the JSON writer must reproduce), hashed
    pj = json.load(open(f"{src}/proof.json"))
    json_sha[f"{kind}_{name[:8]}"] = hashlib.sha256(json.dumps(pj, sort_keys=True, separators=(",", ":")).encode()).hexdigest()

kats = {
    "poseidon_gl_hash_no_pad": {  # gnark-plonky2-verifier/tests/public_inputs_hash_test.go:54-55
        "in": [0, 1, 3736710860384812976],
        "out": [8416658900775745054, 12574228347150446423, 9629056739760131473, 3119289788404190010]},
    "fiat_shamir_test_circuit": {  # gnark-plonky2-verifier/tests/fri_test.go:36-63
        "beta0": 13723612720980423225, "gamma0": 16968338265917648087, "alpha0": 12762198930091358918,
        "zeta0": 6945010638292076810, "fri_alpha0": 1105999839490505715, "fri_beta0_0": 14109475646207874975,
        "pow_response": 57752885224107, "query_index0": 6306766726523865547},
    "poseidon_bn254_permutation": [  # crypto/plonky2_bn128/src/poseidon_bn128.rs:133-180
        {"in": ["0", "0", "0", "0"],
         "out": ["5317387130258456662214331362918410991734007599705406860481038345552731150762",
                 "17768273200467269691696191901389126520069745877826494955630904743826040320364",
                 "19413739268543925182080121099097652227979760828059217876810647045303340666757",
                 "3717738800218482999400886888123026296874264026760636028937972004600663725187"]},
        {"in": ["0", "1", "2", "3"],
         "out": ["6542985608222806190361240322586112750744169038454362455181422643027100751666",
                 "3478427836468552423396868478117894008061261013954248157992395910462939736589",
                 "1904980799580062506738911865015687096398867595589699208837816975692422464009",
                 "11971464497515232077059236682405357499403220967704831154657374522418385384151"]},
        {"in": ["21888242871839275222246405745257275088548364400416034343698204186575808495616"] * 4,
         "out": ["13055670547682322550638362580666986963569035646873545133474324633020685301274",
                 "19087936485076376314486368416882351797015004625427655501762827988254486144933",
                 "10391468779200270580383536396630001155994223659670674913170907401637624483385",
                 "17202557688472898583549180366140168198092766974201433936205272956998081177816"]},
        {"in": ["6542985608222806190361240322586112750744169038454362455181422643027100751666",
                "3478427836468552423396868478117894008061261013954248157992395910462939736589",
                "1904980799580062506738911865015687096398867595589699208837816975692422464009",
                "11971464497515232077059236682405357499403220967704831154657374522418385384151"],
         "out": ["21792249080447013894140672594027696524030291802493510986509431008224624594361",
                 "3536096706123550619294332177231935214243656967137545251021848527424156573335",
                 "14869351042206255711434675256184369368509719143073814271302931417334356905217",
                 "5027523131326906886284185656868809493297314443444919363729302983434650240523"]}],
    "near_proof_json_sha256": json_sha,
}
json.dump(kats, open(f"{OUT}/kats.json", "w"), indent=1)
print("golden vectors written to", OUT)
