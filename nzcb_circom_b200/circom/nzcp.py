"""NZCP templates of /root/reference/circuits/nzcptpl.circom restated on the builder eDSL
(FindCWTClaims :33-145, FindCredSubj :152-227, ReadCredSubj :232-352,
ConstructNullifier :356-434, NZCPPubIdentity :444-655)."""
from .builder import LC, Circuit
from .cbor import (MAJOR_TYPE_INT, MAJOR_TYPE_STRING, copy_string, decode_uint, quin_selector, read_map_length,
                   read_string_length, read_type, skip_value, string_equals)
from .circomlib import bits2num, calculate_total, is_equal, less_than, log2, num2bits, pow_
from .sha2 import sha256_var, sha512_fixed


def _walk_map(c, bytes_, pos, map_len, max_arr, max_map, needle, want_exp):
    """shared body of FindCWTClaims (:69-135) and FindCredSubj (:181-224)"""
    found, exp_pos = [], []
    cur = LC.of(pos)
    for k in range(max_map):
        nxt, typ, v = read_type(c, bytes_, cur)
        value, dnext = decode_uint(c, v, bytes_, nxt)
        is_string = is_equal(c, typ, MAJOR_TYPE_STRING)
        if want_exp:
            is_int = is_equal(c, typ, MAJOR_TYPE_INT)
        skip_pos = dnext + c.mul(value, is_string)
        cur = skip_value(c, bytes_, skip_pos, max_arr)
        is_needle_string = string_equals(c, bytes_, dnext, value, needle)
        if want_exp:
            is4 = is_equal(c, 4, value)
        within = less_than(c, 8, k, map_len)
        is_needle = c.mul(is_string, is_needle_string)
        if want_exp:
            is_exp = c.mul(is_int, is4)
        accepted = c.mul(is_needle, within)
        if want_exp:
            exp_accepted = c.mul(is_exp, within)
        found.append(c.mul(accepted, dnext + value))
        if want_exp:
            exp_pos.append(c.mul(exp_accepted, dnext))
    return calculate_total(found), (calculate_total(exp_pos) if want_exp else None)


def find_cwt_claims(c, bytes_, pos, map_len, max_arr, max_map):
    """FindCWTClaims -- :33-145.  Returns (vcPos, exp)."""
    vc_pos, exp_at = _walk_map(c, bytes_, pos, map_len, max_arr, max_map, [118, 99], True)
    nxt, _typ, v = read_type(c, bytes_, exp_at)                      # :137-140
    exp, _ = decode_uint(c, v, bytes_, nxt)                          # :141-144
    return vc_pos, exp


def find_cred_subj(c, bytes_, pos, map_len, max_arr, max_map):
    """FindCredSubj -- :152-227 (not used by NZCPPubIdentity, :148)"""
    needle = [99, 114, 101, 100, 101, 110, 116, 105, 97, 108, 83, 117, 98, 106, 101, 99, 116]
    needle_pos, _ = _walk_map(c, bytes_, pos, map_len, max_arr, max_map, needle, False)
    return needle_pos


def read_cred_subj(c, bytes_, pos, map_len, max_buffer_len):
    """ReadCredSubj -- :232-352.  Returns ((given, len), (family, len), (dob, len))."""
    MAP_LEN = 3
    max_str = max_buffer_len // MAP_LEN
    GIVEN = [103, 105, 118, 101, 110, 78, 97, 109, 101]
    FAMILY = [102, 97, 109, 105, 108, 121, 78, 97, 109, 101]
    DOB = [100, 111, 98]
    c.assert_eq(map_len, MAP_LEN)                                    # hardcore_assert :261
    is_g, is_f, is_d, copies = [], [], [], []
    cur = LC.of(pos)
    for k in range(MAP_LEN):
        length, nxt = read_string_length(c, bytes_, cur)
        is_g.append(string_equals(c, bytes_, nxt, length, GIVEN))
        is_f.append(string_equals(c, bytes_, nxt, length, FAMILY))
        is_d.append(string_equals(c, bytes_, nxt, length, DOB))
        out, np_, ln = copy_string(c, bytes_, nxt + length, max_str)
        copies.append((out, ln))
        cur = np_

    def route(flags):
        buf = []
        for h in range(max_str):
            buf.append(calculate_total([c.mul(flags[i], copies[i][0][h]) for i in range(MAP_LEN)]))
        buf += [LC()] * (max_buffer_len - max_str)
        ln = calculate_total([c.mul(flags[i], copies[i][1]) for i in range(MAP_LEN)])
        return buf, ln

    return route(is_g), route(is_f), route(is_d)


def construct_nullifier(c, given, given_len, family, family_len, dob, dob_len):
    """ConstructNullifier(MaxBufferLen) -- :356-434.  Returns (result, resultLen)."""
    n = len(given)
    COMMA = 44
    bits = log2(n) + 1
    given_len, family_len, dob_len = LC.of(given_len), LC.of(family_len), LC.of(dob_len)
    result = []
    for k in range(n):
        is_given = less_than(c, bits, k, given_len)
        under_sep1 = less_than(c, bits, k, given_len + 1)
        under_family = less_than(c, bits, k, given_len + 1 + family_len)
        under_sep2 = less_than(c, bits, k, given_len + 1 + family_len + 1)
        g_sel = quin_selector(c, given, k)
        f_sel = quin_selector(c, family, k - given_len - 1)
        d_sel = quin_selector(c, dob, k - given_len - 1 - family_len - 1)
        not_given = 1 + is_given - is_given * 2                         # NOT(in) = 1 + in - 2*in
        is_sep1 = c.mul(under_sep1, not_given)
        is_family = c.mul(under_family, 1 + under_sep1 - under_sep1 * 2)
        is_sep2 = c.mul(under_sep2, 1 + under_family - under_family * 2)
        is_dob = 1 + under_sep2 - under_sep2 * 2
        given_char = c.mul(is_given, g_sel)
        sep1_char = is_sep1 * COMMA
        family_char = c.mul(is_family, f_sel)
        sep2_char = is_sep2 * COMMA
        dob_char = c.mul(is_dob, d_sel)
        result.append(given_char + sep1_char + family_char + sep2_char + dob_char)
    return result, given_len + 1 + family_len + 1 + dob_len


def nzcp_pub_identity(c: Circuit, is_live, max_tbs_bytes, max_arr_vc, max_map_vc, max_arr_cs, max_map_cs):
    """NZCPPubIdentity -- :444-655; main component, declares its own I/O."""
    CHUNK_BITS, BYTE_BITS, TS_BYTES = 248, 8, 4
    CHUNK_BYTES = CHUNK_BITS // BYTE_BITS
    VC_OFFSET, CS_MAP_LEN, NULL_BYTES = 171, 3, 64
    DATA_LEN = 20 * BYTE_BITS
    claims_skip = 30 if is_live else 27
    max_bits = max_tbs_bytes * 8
    block_space = 3
    block_count = pow_(2, block_space)
    max_sha_bits = 512 * block_count
    assert max_bits <= max_sha_bits

    out = c.output("out", 3)
    tbs = c.input("toBeSigned", max_bits)
    tbs_len = c.input("toBeSignedLen")
    data = c.input("data", DATA_LEN)

    for i in range(max_bits):                                            # :493-496
        c.assert_zero(tbs[i] * (tbs[i] - 1))
    lte = less_than(c, log2(max_tbs_bytes + 1) + 1, tbs_len, max_tbs_bytes + 1)   # :500-505
    c.assert_eq(lte, 1)

    sha_in = list(tbs) + [LC()] * (max_sha_bits - max_bits)             # :509-516
    tbs_hash = sha256_var(c, sha_in, tbs_len * 8, block_space)

    ToBeSigned = []                                                      # :521-533
    lbits = log2(max_tbs_bytes) + 1
    for k in range(max_tbs_bytes):
        b2n = bits2num([tbs[k * 8 + (7 - i)] for i in range(8)])
        lt = less_than(c, lbits, k, tbs_len)
        ToBeSigned.append(c.mul(b2n, lt))

    map_len, nxt = read_map_length(c, ToBeSigned, claims_skip)          # :535-537
    vc_pos, exp = find_cwt_claims(c, ToBeSigned, nxt, map_len, max_arr_vc, max_map_vc)   # :541-545
    (given, gl), (family, fl), (dob, dl) = read_cred_subj(c, ToBeSigned, VC_OFFSET + vc_pos, CS_MAP_LEN, NULL_BYTES)
    result, _rlen = construct_nullifier(c, given, gl, family, fl, dob, dl)             # :554-560

    null_bits = [None] * (NULL_BYTES * 8)                                # :563-571
    for k in range(NULL_BYTES):
        nb = num2bits(c, result[k], 8)
        for j in range(8):
            null_bits[k * 8 + (7 - j)] = nb[j]
    h512 = sha512_fixed(c, null_bits)                                    # :577-580

    exp_bits = num2bits(c, exp, TS_BYTES * BYTE_BITS)                    # :583-584
    o = [[LC()] * CHUNK_BITS for _ in range(3)]
    for k in range(CHUNK_BYTES):                                         # :595-600
        b = CHUNK_BYTES - 1 - k
        for i in range(BYTE_BITS):
            o[0][b * 8 + (7 - i)] = h512[k * 8 + i]
    for k in range(8 // BYTE_BITS):                                      # :601-606
        b = CHUNK_BYTES - 1 - k
        for i in range(BYTE_BITS):
            o[1][b * 8 + (7 - i)] = h512[CHUNK_BITS + k * 8 + i]
    for k in range(1, CHUNK_BYTES):                                      # :609-614
        b = CHUNK_BYTES - 1 - k
        for i in range(BYTE_BITS):
            o[1][b * 8 + (7 - i)] = tbs_hash[(k * 8 + i) - 8]
    for k in range(16 // BYTE_BITS):                                     # :615-620
        b = CHUNK_BYTES - 1 - k
        for i in range(BYTE_BITS):
            o[2][b * 8 + (7 - i)] = tbs_hash[CHUNK_BITS + (k * 8 + i) - 8]
    idx = 0                                                              # :625-633
    for k in range(2, 2 + TS_BYTES):
        b = CHUNK_BYTES - 1 - k
        d = TS_BYTES - 1 - idx
        for i in range(BYTE_BITS):
            o[2][b * 8 + i] = exp_bits[d * 8 + i]
        idx += 1
    idx = 0                                                              # :636-650
    data_bytes = DATA_LEN // BYTE_BITS
    for k in range(2 + TS_BYTES, CHUNK_BYTES):
        b = CHUNK_BYTES - 1 - k
        for i in range(BYTE_BITS):
            if idx < data_bytes:
                d = data_bytes - 1 - idx
                o[2][b * 8 + i] = data[d * 8 + i]
        idx += 1
    for j in range(3):                                                   # :652-654
        c.assign_output(out[j], bits2num(o[j]))
    return c
