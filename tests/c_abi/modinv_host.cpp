// Host build of csrc/modinv.cuh (the same source the device compiles): reads "modulus_hex value_hex" lines from
// stdin (64 hex digits each, big-endian), prints value^-1 mod modulus.  tests/test_modinv_host.py compares with Python.
#include <stdio.h>
#include <string.h>
#include "../../cudabulletproof_b200/csrc/modinv.cuh"

static int parse(const char* hex, uint32_t (&w)[8]) {
    if (strlen(hex) != 64) return 0;
    for (int i = 0; i < 8; i++) {
        unsigned v;
        char buf[9];
        memcpy(buf, hex + 56 - 8 * i, 8);
        buf[8] = 0;
        if (sscanf(buf, "%x", &v) != 1) return 0;
        w[i] = v;
    }
    return 1;
}

int main() {
    char a[128], b[128];
    while (scanf("%100s %100s", a, b) == 2) {
        uint32_t m[8], x[8], r[8];
        if (!parse(a, m) || !parse(b, x)) return 2;
        cbp::ModInfo mi = cbp::modinfo_from_words(m);
        cbp::modinv_words(r, x, mi);
        for (int i = 7; i >= 0; i--) printf("%08x", r[i]);
        printf("\n");
    }
    return 0;
}
