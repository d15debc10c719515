// Host-side input pipeline helper: parse a libsvm text file (`label id:val id:val ...`, data_loader.py:12-47) straight
// into the dense [N, F] arrays the kernels consume (ids int32, values fp32, labels fp32).  The reference parses every
// line with Python string ops (~1e5 lines/s); Criteo is 45 M lines.  Lines that do not hold exactly nfield well-formed
// pairs are skipped and counted, like the reference's try/except around each line (data_loader.py:37-44).
#include <errno.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

using namespace armnet;

extern "C" int armnet_libsvm_count_lines(const char *path, int64_t *n_lines) {
    if (!path || !n_lines) {
        set_error("armnet_libsvm_count_lines: null pointer");
        return ARMNET_ERR_NULL;
    }
    FILE *f = fopen(path, "rb");
    if (!f) {
        set_error("armnet_libsvm_count_lines: cannot open %s: %s", path, strerror(errno));
        return ARMNET_ERR_SHAPE;
    }
    static const size_t kBuf = 1 << 20;
    char *buf = (char *)malloc(kBuf);
    int64_t n = 0;
    size_t got;
    char last = '\n';
    while ((got = fread(buf, 1, kBuf, f)) > 0) {
        for (size_t i = 0; i < got; ++i) n += buf[i] == '\n';
        last = buf[got - 1];
    }
    if (last != '\n') ++n;  // final line without a newline
    free(buf);
    fclose(f);
    *n_lines = n;
    return ARMNET_OK;
}

extern "C" int armnet_libsvm_parse(const char *path, int nfield, int64_t capacity, int32_t *ids, float *values,
                                   float *labels, int64_t *n_rows, int64_t *n_skipped) {
    if (!path || !ids || !values || !labels || !n_rows) {
        set_error("armnet_libsvm_parse: null pointer");
        return ARMNET_ERR_NULL;
    }
    if (nfield <= 0 || capacity < 0) {
        set_error("armnet_libsvm_parse: bad sizes nfield=%d capacity=%lld", nfield, (long long)capacity);
        return ARMNET_ERR_SHAPE;
    }
    FILE *f = fopen(path, "rb");
    if (!f) {
        set_error("armnet_libsvm_parse: cannot open %s: %s", path, strerror(errno));
        return ARMNET_ERR_SHAPE;
    }
    char *line = nullptr;
    size_t cap = 0;
    int64_t rows = 0, bad = 0;
    ssize_t len;
    while ((len = getline(&line, &cap, f)) >= 0) {
        char *p = line;
        while (*p == ' ' || *p == '\t') ++p;
        if (*p == '\n' || *p == '\r' || *p == '\0') {
            if (len > 0) ++bad;  // the reference reports blank lines as malformed too
            continue;
        }
        char *end;
        errno = 0;
        const float y = strtof(p, &end);
        bool ok = end != p && (*end == ' ' || *end == '\t');
        int nf = 0;
        int32_t *irow = ids + rows * nfield;
        float *vrow = values + rows * nfield;
        p = end;
        while (ok) {
            while (*p == ' ' || *p == '\t') ++p;
            if (*p == '\n' || *p == '\r' || *p == '\0') break;
            const long long id = strtoll(p, &end, 10);
            if (end == p || *end != ':' || id < 0 || id > 2147483647LL) {
                ok = false;
                break;
            }
            p = end + 1;
            const float v = strtof(p, &end);
            if (end == p) {
                ok = false;
                break;
            }
            if (nf < nfield && rows < capacity) {
                irow[nf] = (int32_t)id;
                vrow[nf] = v;
            }
            ++nf;
            p = end;
        }
        if (!ok || nf != nfield) {
            ++bad;
            continue;
        }
        if (rows >= capacity) {
            free(line);
            fclose(f);
            set_error("armnet_libsvm_parse: more than %lld well-formed rows in %s", (long long)capacity, path);
            return ARMNET_ERR_SHAPE;
        }
        labels[rows] = y;
        ++rows;
    }
    free(line);
    fclose(f);
    *n_rows = rows;
    if (n_skipped) *n_skipped = bad;
    return ARMNET_OK;
}
