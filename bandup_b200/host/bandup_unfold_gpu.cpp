// bandup_unfold_gpu.cpp -- BandUP_gpu.x: the host side of `bandup unfold` above the C ABI, in C++.
//
// The reference's unfold driver is compiled Fortran (program BandUP_main, src/main_BandUP.f90);
// no Fortran compiler exists in this build image, so the native host that exercises
// libbandup_gpu.so is this C++ mirror of it (the ISO_C_BINDING shim a BandUP maintainer would use
// instead is fortran/bandup_gpu_mod.f90).  It takes the reference's own
// command-line flags (cla_wrappers_mod.f90:152-341) and input files and writes the reference's
// output files, for the no-symmetry semantics (-no_symm_sckpts -no_symm_avg: identity operation
// only, one needed direction of weight 1; band_unfolding_routines_mod.f90:192-204,
// symmetry_mod.f90:1009-1023) -- symmetry expansion/averaging needs spglib and stays with BandUP.
//
// Flow (main_BandUP.f90:52-198):
//   read cells -> verify_commens -> read pc-kpt path -> energy grid -> for every SC-kpt of the
//   WAVECAR with folding pc-kpts: raw band records (parallel pread into page-locked memory, the
//   next k-point read while this one uploads) -> bandup_gpu_submit_kpt_vasp -> fetch ->
//   [rho, Pauli vectors] -> write unfolded_EBS_not-symmetry_averaged.dat [+ rho file].
// Every number comes from the GPU library; nothing is computed on the host.
#include <fcntl.h>
#include <sys/file.h>
#include <unistd.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <fstream>
#include <future>
#include <map>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "bandup_gpu.h"

namespace {

using Vec3 = double[3];
struct Mat3 { double m[9]; };  // row-major, row i = vector i

[[noreturn]] void die(const std::string& msg) {
  std::fprintf(stderr, "%s\nStopping now.\n", msg.c_str());
  std::exit(1);
}

void gpu_check(int handle, int rc, const char* what) {
  if (rc != BANDUP_GPU_OK) die(std::string("ERROR (") + what + "): " + bandup_gpu_last_error(handle));
}

const double kTwoPi = 2.0 * (4.0 * std::atan(1.0));

void cross(const double* a, const double* b, double* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

Mat3 rec_latt(const Mat3& a) {  // crystals_mod.f90:186-202
  Mat3 b;
  double c[3];
  cross(a.m + 3, a.m + 6, c);
  const double v = std::fabs(a.m[0] * c[0] + a.m[1] * c[1] + a.m[2] * c[2]);
  for (int j = 0; j < 3; ++j) b.m[j] = kTwoPi * c[j] / v;
  cross(a.m + 6, a.m + 0, c);
  for (int j = 0; j < 3; ++j) b.m[3 + j] = kTwoPi * c[j] / v;
  cross(a.m + 0, a.m + 3, c);
  for (int j = 0; j < 3; ++j) b.m[6 + j] = kTwoPi * c[j] / v;
  return b;
}

bool inverse(const Mat3& a, Mat3* inv) {
  const double* m = a.m;
  const double c0 = m[4] * m[8] - m[5] * m[7], c1 = -(m[3] * m[8] - m[5] * m[6]), c2 = m[3] * m[7] - m[4] * m[6];
  const double det = m[0] * c0 + m[1] * c1 + m[2] * c2;
  if (std::fabs(det) < 1e-14) return false;
  double* r = inv->m;
  r[0] = c0 / det; r[3] = c1 / det; r[6] = c2 / det;
  r[1] = -(m[1] * m[8] - m[2] * m[7]) / det; r[4] = (m[0] * m[8] - m[2] * m[6]) / det; r[7] = -(m[0] * m[7] - m[1] * m[6]) / det;
  r[2] = (m[1] * m[5] - m[2] * m[4]) / det; r[5] = -(m[0] * m[5] - m[2] * m[3]) / det; r[8] = (m[0] * m[4] - m[1] * m[3]) / det;
  return true;
}

void vec_mat(const double* v, const Mat3& a, double* o) {  // row vector times matrix
  for (int j = 0; j < 3; ++j) o[j] = v[0] * a.m[j] + v[1] * a.m[3 + j] + v[2] * a.m[6 + j];
}

void to_fortran(const Mat3& a, double* f) {  // element (i,j) at [j*3+i]
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) f[3 * j + i] = a.m[3 * i + j];
}

std::string strip_comment(std::string s) {
  for (char c : {'!', '#'}) {
    const size_t p = s.find(c);
    if (p != std::string::npos) s = s.substr(0, p);
  }
  return s;
}

std::vector<std::string> read_lines(const std::string& path) {
  std::ifstream f(path);
  if (!f) die("ERROR: File \"" + path + "\" not found.");
  std::vector<std::string> out;
  std::string ln;
  while (std::getline(f, ln)) out.push_back(ln);
  return out;
}

bool blank(const std::string& s) { return s.find_first_not_of(" \t\r\n") == std::string::npos; }

// POSCAR-style lattice (read_vasp_files_mod.f90:63-200): comment / scale / three vectors
Mat3 read_lattice(const std::string& path) {
  const auto ln = read_lines(path);
  if (ln.size() < 5) die("ERROR reading lattice file " + path);
  const double scale = std::atof(ln[1].c_str());
  Mat3 a;
  for (int i = 0; i < 3; ++i) {
    std::istringstream is(strip_comment(ln[2 + i]));
    if (!(is >> a.m[3 * i] >> a.m[3 * i + 1] >> a.m[3 * i + 2])) die("ERROR reading lattice vectors in " + path);
    for (int j = 0; j < 3; ++j) a.m[3 * i + j] *= scale;
  }
  return a;
}

struct KPath {
  std::vector<std::vector<std::array<double, 3>>> dirs;
  double zero = 0.0;
  size_t total() const {
    size_t n = 0;
    for (auto& d : dirs) n += d.size();
    return n;
  }
};

// read_pckpts_selected_by_user (io_routines_mod.f90:981-1211) + kpts_line (lists_and_seqs_mod.f90:219-241)
KPath read_pckpts(const std::string& path, const Mat3& b_pc) {
  const auto ln = read_lines(path);
  if (ln.size() < 6) die("ERROR reading input pc-kpts file. Please check the format.");
  std::vector<int> counts;
  {
    std::istringstream is(ln[1]);
    int n;
    while (is >> n) counts.push_back(n);
    if (counts.empty()) die("ERROR: automatic scan for pc-kpts is not supported by BandUP_gpu.x; give the numbers of k-points.");
  }
  KPath kp;
  {
    std::istringstream is(ln[2]);
    std::string tag;
    is >> tag;
    if (!(is >> kp.zero)) kp.zero = 0.0;
  }
  bool cartesian = false;
  double a0 = 0.0;
  {
    std::istringstream is(ln[3]);
    std::string tag;
    is >> tag;
    cartesian = !tag.empty() && (tag[0] == 'C' || tag[0] == 'c');
    if (!(is >> a0)) a0 = 0.0;
    if (cartesian && a0 == 0.0) {
      std::istringstream first(ln[0]);
      if (!(first >> a0) || a0 == 0.0)
        die("ERROR: You have selected cartesian coordinates in your input k-points file, but you have not passed a scaling parameter \"a0\".");
    }
  }
  std::vector<std::array<double, 3>> block;
  std::vector<std::pair<std::array<double, 3>, std::array<double, 3>>> pairs;
  for (size_t i = 4; i <= ln.size(); ++i) {
    const std::string s = i < ln.size() ? strip_comment(ln[i]) : std::string();
    if (blank(s)) {
      if (block.size() == 2) pairs.push_back({block[0], block[1]});
      else if (!block.empty()) die("ERROR reading input pc-kpts file. Please check the format.");
      block.clear();
    } else {
      std::array<double, 3> v;
      std::istringstream is(s);
      if (!(is >> v[0] >> v[1] >> v[2])) die("ERROR reading input pc-kpts file.");
      block.push_back(v);
    }
  }
  if (pairs.empty()) die("ERROR reading input pc-kpts file. Please check the format.");
  while (counts.size() < pairs.size()) counts.push_back(counts[0]);
  for (size_t d = 0; d < pairs.size(); ++d) {
    double ks[3], ke[3];
    if (cartesian) {
      for (int j = 0; j < 3; ++j) {
        ks[j] = kTwoPi * pairs[d].first[j] / a0;
        ke[j] = kTwoPi * pairs[d].second[j] / a0;
      }
    } else {
      vec_mat(pairs[d].first.data(), b_pc, ks);
      vec_mat(pairs[d].second.data(), b_pc, ke);
    }
    int nk = counts[d];
    const double len = std::sqrt((ks[0] - ke[0]) * (ks[0] - ke[0]) + (ks[1] - ke[1]) * (ks[1] - ke[1]) + (ks[2] - ke[2]) * (ks[2] - ke[2]));
    if (len < 2.220446049250313e-16) nk = 1;
    std::vector<std::array<double, 3>> line((size_t)nk);
    for (int i = 0; i < nk; ++i)
      for (int j = 0; j < 3; ++j) {
        const double step = nk > 1 ? (ke[j] - ks[j]) / (double)(nk - 1) : 0.0;
        line[(size_t)i][j] = ks[j] + (double)i * step;
      }
    kp.dirs.push_back(line);
  }
  return kp;
}

// read_energy_info_for_band_search (io_routines_mod.f90:334-388)
void read_energy_info(const std::string& path, double* ef, double* e0, double* e1, double* de) {
  const auto ln = read_lines(path);
  std::vector<std::string> vals;
  for (auto& l : ln) {
    std::string s = l;
    const size_t p = s.find_first_not_of(" \t");
    if (p == std::string::npos || s[p] == '#') continue;
    std::istringstream is(strip_comment(s));
    std::string tok;
    if (is >> tok) vals.push_back(tok);
    if (vals.size() == 4) break;
  }
  if (vals.size() < 4) die("ERROR reading " + path + ": need E_Fermi, Emin-E_F, Emax-E_F and dE");
  *ef = std::atof(vals[0].c_str());
  *e0 = std::atof(vals[1].c_str()) + *ef;
  *e1 = std::atof(vals[2].c_str()) + *ef;
  if (*e1 < *e0) std::swap(*e0, *e1);
  char* end = nullptr;
  const double d = std::strtod(vals[3].c_str(), &end);
  *de = (end != vals[3].c_str()) ? std::fabs(d) : std::fabs(0.4e-2 * (*e1 - *e0));
}

// ---- WAVECAR (read_wavecar_mod.f90:307-587) --------------------------------------------------
struct Wavecar {
  int fd = -1;
  int64_t nrecl = 0;
  int nspin = 0, nkpts = 0, nbands = 0;
  double encut = 0.0;
  Mat3 A;

  void pread_all(void* dst, size_t n, int64_t off) const {
    char* p = static_cast<char*>(dst);
    size_t done = 0;
    while (done < n) {
      const ssize_t k = ::pread(fd, p + done, n - done, off + (int64_t)done);
      if (k <= 0) die("ERROR (read_wavecar): short read");
      done += (size_t)k;
    }
  }
  void open(const std::string& path) {
    fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) die("ERROR: File \"" + path + "\" not found.");
    double h[3];
    pread_all(h, 24, 0);
    nrecl = std::llround(h[0]);
    nspin = (int)std::lround(h[1]);
    if (std::lround(h[2]) != 45200)
      die("ERROR (read_wavecar module): You seem to be using a double-precision WAVECAR (WAVECAR_double); only complex64 coefficients are supported.");
    double r2[12];
    pread_all(r2, 96, nrecl);
    nkpts = (int)std::lround(r2[0]);
    nbands = (int)std::lround(r2[1]);
    encut = r2[2];
    if (nkpts == 0) die("ERROR (read_wavecar): Unexpected file format");
    std::memcpy(A.m, r2 + 3, 72);
  }
  int64_t header_record(int ikpt, int ispin) const {  // 0-based record of the k-point header (:409-414)
    return 2 + (int64_t)(ispin - 1) * nkpts * (1 + nbands) + (int64_t)(ikpt - 1) * (1 + nbands);
  }
  void kpt_header(int ikpt, int ispin, int* nplane, double* kfrac, std::vector<double>* ener) const {
    std::vector<double> h(4 + 3 * (size_t)nbands);
    pread_all(h.data(), h.size() * 8, header_record(ikpt, ispin) * nrecl);
    *nplane = (int)std::lround(h[0]);
    for (int j = 0; j < 3; ++j) kfrac[j] = h[1 + j];
    ener->resize((size_t)nbands);
    for (int b = 0; b < nbands; ++b) (*ener)[(size_t)b] = h[4 + 3 * (size_t)b];
  }
  // the n_bands raw coefficient records of a k-point: contiguous on disk, read by several threads
  void read_band_records(int ikpt, int ispin, void* dst, int n_threads) const {
    const size_t n = (size_t)nbands * (size_t)nrecl;
    const int64_t base = (header_record(ikpt, ispin) + 1) * nrecl;
    int nt = std::max(1, std::min<int>(n_threads, (int)(n / (4u << 20))));
    if (nt == 1) {
      pread_all(dst, n, base);
      return;
    }
    size_t step = (n + nt - 1) / nt;
    step += (4096 - step % 4096) % 4096;
    std::vector<std::thread> th;
    for (size_t lo = 0; lo < n; lo += step) {
      const size_t hi = std::min(lo + step, n);
      th.emplace_back([=] { pread_all(static_cast<char*>(dst) + lo, hi - lo, base + (int64_t)lo); });
    }
    for (auto& t : th) t.join();
  }
};

std::string es10_3(double x) {  // Fortran ES10.3; three-digit exponents drop the letter (d.ddd+eee)
  char buf[40];
  std::snprintf(buf, sizeof buf, "%.3E", x);
  std::string s = buf;
  const size_t e = s.find('E');
  const int ex = std::atoi(s.c_str() + e + 1);
  if (ex > 99 || ex < -99) {
    std::snprintf(buf, sizeof buf, "%s%+04d", s.substr(0, e).c_str(), ex);
    s = buf;
  }
  if (s.size() > 10) return std::string(10, '*');
  return std::string(10 - s.size(), ' ') + s;
}

struct Args {
  std::map<std::string, std::string> kv;
  std::map<std::string, bool> flag;
  std::string get(const std::string& k, const std::string& def) const {
    auto it = kv.find(k);
    return it == kv.end() ? def : it->second;
  }
  bool has(const std::string& k) const { return flag.count(k) != 0; }
  bool has_value(const std::string& k) const { return kv.count(k) != 0; }
};

Args parse_args(int argc, char** argv) {
  // value flags and switches of cla_wrappers_mod.f90:152-341 that matter on this path
  const std::vector<std::string> with_value = {"-wf_file", "-pc_file", "-sc_file", "-pckpts_file", "-energy_file",
      "-out_file_nosymm", "-out_file_symm", "-output_file_only_user_selec_direcs_unf_dens_op",
      "-output_file_symm_averaged_unf_dens_op", "-spin_channel", "-n_sckpts_to_skip", "-normal_to_proj_plane",
      "-origin_for_spin_proj_cart", "-origin_for_spin_proj_rec", "-device", "-read_threads", "-outdir", "-saxis", "-n_ranks", "-rank",
      "-nccl_id_file", "-ticket_file"};
  const std::vector<std::string> switches = {"-write_unf_dens_op", "-no_symm_sckpts", "-no_symm_avg", "-dont_unfold",
      "-continue_if_not_commensurate", "--continue_if_npw_smaller_than_expected", "-skip_propose_pc_for_given_pc",
      "-skip_propose_pc_for_given_sc"};
  Args a;
  for (int i = 1; i < argc; ++i) {
    const std::string k = argv[i];
    if (std::find(with_value.begin(), with_value.end(), k) != with_value.end()) {
      if (i + 1 >= argc) die("ERROR: flag " + k + " needs a value");
      a.kv[k] = argv[++i];
    } else if (std::find(switches.begin(), switches.end(), k) != switches.end()) {
      a.flag[k] = true;
    } else if (k == "-qe" || k == "-abinit" || k == "-castep") {
      die("ERROR: BandUP_gpu.x reads VASP WAVECAR files; for " + k + " use BandUP.x with the bandup_gpu module (INTEGRATION.md).");
    } else {
      die("ERROR: unknown command-line option " + k);
    }
  }
  return a;
}

void parse_vec3(const std::string& s, double* v) {
  std::string t = s;
  std::replace(t.begin(), t.end(), ',', ' ');
  std::istringstream is(t);
  if (!(is >> v[0] >> v[1] >> v[2])) die("ERROR: could not read a 3-vector from \"" + s + "\"");
}

}  // namespace

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// print_time (io_routines_mod.f90:1885-1920): message, seconds, share of the total
void print_time(const char* message, double t, double total) {
  if (t == total)
    std::printf("%s%7.2fs.\n", message, t);
  else if (t >= 0.1 || (total > 0.0 && 100.0 * t / total >= 1.0))
    std::printf("%s%7.2fs (%.1f%%).\n", message, t, total > 0.0 ? 100.0 * t / total : 0.0);
}

int main(int argc, char** argv) {
  const double t_start = now_s();
  double t_read = 0.0, t_write = 0.0;  // t_read is only touched by the one read in flight
  const Args args = parse_args(argc, argv);
  const std::string wf_file = args.get("-wf_file", "WAVECAR");
  const std::string out_file = args.get("-out_file_nosymm", "unfolded_EBS_not-symmetry_averaged.dat");
  const std::string udo_file = args.get("-output_file_only_user_selec_direcs_unf_dens_op",
                                        "unfolding_density_operator_not-symm_avgd.dat");
  const int spin_channel = std::atoi(args.get("-spin_channel", "1").c_str());
  const int n_skip = std::atoi(args.get("-n_sckpts_to_skip", "0").c_str());
  const int read_threads = std::atoi(args.get("-read_threads", "16").c_str());
  const bool write_udo = args.has("-write_unf_dens_op");
  if (!args.has("-no_symm_sckpts") || !args.has("-no_symm_avg"))
    std::printf("NOTICE: BandUP_gpu.x works without symmetry (as if -no_symm_sckpts -no_symm_avg had been given).\n");

  // one process per GPU: rank r of n_ranks unfolds a contiguous block of the SC-kpts and the
  // delta_N rows are gathered with NCCL (the 128-byte id travels through -nccl_id_file)
  const int n_ranks = std::atoi(args.get("-n_ranks", "1").c_str());
  const int rank = std::atoi(args.get("-rank", "0").c_str());
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks) die("ERROR: bad -rank / -n_ranks");

  int h = -1;
  gpu_check(-1, bandup_gpu_init(std::atoi(args.get("-device", std::to_string(rank)).c_str()), &h), "bandup_gpu_init");
  gpu_check(h, bandup_gpu_set_profiling(h, 1), "set_profiling");  // per-kernel times for the final report
  if (n_ranks > 1) {
    const std::string idf = args.get("-nccl_id_file", "bandup_gpu_nccl_id");
    char id[BANDUP_GPU_COMM_ID_BYTES];
    if (rank == 0) {
      gpu_check(h, bandup_gpu_comm_unique_id(id), "comm_unique_id");
      std::ofstream o(idf + ".tmp", std::ios::binary);
      o.write(id, sizeof id);
      o.close();
      std::rename((idf + ".tmp").c_str(), idf.c_str());
    } else {
      for (int tries = 0;; ++tries) {
        std::ifstream in(idf, std::ios::binary);
        if (in && in.read(id, sizeof id)) break;
        if (tries > 1200) die("ERROR: no NCCL id from rank 0 in " + idf);
        std::this_thread::sleep_for(std::chrono::milliseconds(50));
      }
    }
    gpu_check(h, bandup_gpu_comm_init(h, n_ranks, rank, id), "comm_init");
  }

  // cells and commensurability (main_BandUP.f90:57-82)
  Wavecar wc;
  wc.open(wf_file);
  const Mat3 a_pc = read_lattice(args.get("-pc_file", "prim_cell_lattice.in"));
  const Mat3 A_sc = read_lattice(args.get("-sc_file", "supercell_lattice.in"));
  double worst = 0.0, scale = 0.0;
  for (int i = 0; i < 9; ++i) {
    worst = std::max(worst, std::fabs(A_sc.m[i] - wc.A.m[i]));
    scale = std::max(scale, std::fabs(A_sc.m[i]));
  }
  if (worst > 1e-4 * scale) die("ERROR: the supercell of " + args.get("-sc_file", "supercell_lattice.in") + " is not the one of the wavefunction file.");
  const Mat3 b_pc = rec_latt(a_pc), B_sc = rec_latt(wc.A);
  double Bf[9], bf[9];
  int32_t M[9];
  to_fortran(B_sc, Bf);
  to_fortran(b_pc, bf);
  {
    const int rc = bandup_gpu_set_cells(h, Bf, bf, 0.0, M);
    if (rc == BANDUP_GPU_ERR_NOT_COMMENSURATE && !args.has("-continue_if_not_commensurate"))
      die(std::string("ERROR: The SC and the pc are not commensurate! ") + bandup_gpu_last_error(h));
    gpu_check(h, rc, "verify_commens");
  }
  if (args.has("-dont_unfold")) {
    std::printf("    **** WARNING: You are using the option \"-dont_unfold\". You are NOT performing unfolding!! ****\n");
    gpu_check(h, bandup_gpu_set_perform_unfold(h, 0), "set_perform_unfold");
  }

  // pc k-points and energy grid (:84-124)
  const KPath kp = read_pckpts(args.get("-pckpts_file", "KPOINTS_prim_cell.in"), b_pc);
  double ef, e0, e1, de;
  read_energy_info(args.get("-energy_file", "energy_info.in"), &ef, &e0, &e1, &de);
  int32_t nener = 0;
  gpu_check(h, bandup_gpu_set_grid(h, e0, e1, de, -1.0, &nener), "set_grid");
  std::vector<double> grid((size_t)nener);
  gpu_check(h, bandup_gpu_get_grid(h, grid.data()), "get_grid");

  // geometric unfolding relations without symmetry: k folds onto K iff k - K is an SC reciprocal vector
  std::vector<std::array<double, 3>> pck;
  for (auto& d : kp.dirs) pck.insert(pck.end(), d.begin(), d.end());
  const size_t n_pck = pck.size();
  Mat3 Binv;  // only for the fractional coordinates printed in the rho file
  if (!inverse(B_sc, &Binv)) die("ERROR: singular SC reciprocal lattice");
  std::vector<std::array<double, 3>> Kfrac((size_t)wc.nkpts), Kcart((size_t)wc.nkpts);
  std::vector<int> nplane((size_t)wc.nkpts);
  std::vector<std::vector<double>> ener((size_t)wc.nkpts);
  for (int i = 0; i < wc.nkpts; ++i) {
    wc.kpt_header(i + 1, spin_channel, &nplane[(size_t)i], Kfrac[(size_t)i].data(), &ener[(size_t)i]);
    vec_mat(Kfrac[(size_t)i].data(), B_sc, Kcart[(size_t)i].data());
  }
  // get_geom_unfolding_relations (band_unfolding_routines_mod.f90:430-483) on the device; without
  // symmetry every SC-kpt is its own only equivalent point
  std::vector<std::vector<int>> folds((size_t)wc.nkpts);
  std::vector<int> sck_of((size_t)n_pck, -1);
  {
    std::vector<double> pk(n_pck * 3), eq((size_t)wc.nkpts * 3);
    std::vector<int32_t> off((size_t)wc.nkpts + 1), fs(n_pck);
    for (size_t i = 0; i < n_pck; ++i)
      for (int j = 0; j < 3; ++j) pk[3 * i + (size_t)j] = pck[i][(size_t)j];
    for (int i = 0; i < wc.nkpts; ++i) {
      off[(size_t)i] = i;
      for (int j = 0; j < 3; ++j) eq[3 * (size_t)i + (size_t)j] = Kcart[(size_t)i][(size_t)j];
    }
    off[(size_t)wc.nkpts] = wc.nkpts;
    double tol_used = 0.0;
    const int rc = bandup_gpu_fold_map(h, (int)n_pck, pk.data(), wc.nkpts, off.data(), eq.data(), n_skip, fs.data(),
                                       nullptr, &tol_used);
    if (rc == BANDUP_GPU_ERR_FOLDING_VECTOR) {  // main_BandUP.f90:114
      for (size_t ik = 0; ik < n_pck; ++ik)
        if (fs[ik] < 0)
          die("ERROR: pc-kpt #" + std::to_string(ik + 1) + " does not fold onto any SC-kpt of the wavefunction file.");
    }
    gpu_check(h, rc, "get_geom_unfolding_relations");
    if (tol_used > 1.0000001e-5)
      std::printf("WARNING (get_geom_unfolding_relations): the tolerance for vector equality had to be raised to %.3e.\n", tol_used);
    for (size_t ik = 0; ik < n_pck; ++ik) {
      sck_of[ik] = fs[ik];
      folds[(size_t)fs[ik]].push_back((int)ik);
    }
  }
  std::vector<int> used_all, used, mine_static;
  for (int i = 0; i < wc.nkpts; ++i)
    if (!folds[(size_t)i].empty()) used_all.push_back(i);
  {  // this rank's contiguous block, balanced by n_bands * nplane (main_BandUP.f90:133 becomes first..last)
    double total = 0.0;
    for (int i : used_all) total += (double)nplane[(size_t)i];
    double acc = 0.0;
    for (int i : used_all) {
      const double mid = acc + 0.5 * (double)nplane[(size_t)i];
      const int owner = std::min(n_ranks - 1, (int)(mid / total * n_ranks));
      if (owner == rank) mine_static.push_back(i);
      acc += (double)nplane[(size_t)i];
    }
  }
  // -ticket_file F: instead of the static blocks, every rank takes the next SC-kpt from a counter
  // in F (flock + fetch-and-add; F must not exist when the job starts), so that a rank behind a
  // slower host link simply unfolds fewer k-points.  The delta_N rows are gathered by row id.
  const std::string ticket_file = args.get("-ticket_file", "");
  size_t static_pos = 0;
  auto next_used = [&]() -> bool {  // appends this rank's next SC-kpt to `used`
    if (ticket_file.empty()) {
      if (static_pos >= mine_static.size()) return false;
      used.push_back(mine_static[static_pos++]);
      return true;
    }
    const int fd = ::open(ticket_file.c_str(), O_RDWR | O_CREAT, 0644);
    if (fd < 0) die("ERROR: cannot open -ticket_file " + ticket_file);
    int64_t v = 0;
    if (::flock(fd, LOCK_EX) != 0) die("ERROR: cannot lock -ticket_file " + ticket_file);
    if (::pread(fd, &v, sizeof(v), 0) != (ssize_t)sizeof(v)) v = 0;
    const int64_t nv = v + 1;
    if (::pwrite(fd, &nv, sizeof(nv), 0) != (ssize_t)sizeof(nv)) die("ERROR: cannot write -ticket_file " + ticket_file);
    ::flock(fd, LOCK_UN);
    ::close(fd);
    if (v >= (int64_t)used_all.size()) return false;
    used.push_back(used_all[(size_t)v]);
    return true;
  };

  // MAIN LOOP (:131-177), double buffered: records of k-point j+1 are read while j uploads/computes
  const size_t block = (size_t)wc.nbands * (size_t)wc.nrecl;
  const int n_buf = (int)std::min<size_t>(4, std::max<size_t>(1, ticket_file.empty() ? mine_static.size() : used_all.size()));
  std::vector<void*> bufs((size_t)n_buf, nullptr);
  for (auto& b : bufs) gpu_check(h, bandup_gpu_pinned_alloc(block, &b), "pinned_alloc");
  std::vector<double> dN_all(n_pck * (size_t)nener, 0.0), sigma_all;
  bool spinor = false;
  struct RhoRows { std::vector<int32_t> ie, m1, m2; std::vector<float> re, im; };
  std::vector<RhoRows> rho_of(n_pck);
  const int depth = bandup_gpu_pipeline_depth(h);
  std::vector<int> inflight;  // job indices
  size_t n_parsed = 0;

  auto drain = [&](int j) {
    const int i = used[(size_t)j];
    const auto& rows = folds[(size_t)i];
    const int nf = (int)rows.size();
    std::vector<double> dN((size_t)nf * (size_t)nener);
    std::vector<int32_t> ns((size_t)nf);
    gpu_check(h, bandup_gpu_fetch_kpt(h, j + 1, nullptr, dN.data(), ns.data(), nullptr, nullptr, 0), "fetch_kpt");
    for (int f = 0; f < nf; ++f) {
      if (ns[(size_t)f] == 0) {  // print_message_pckpt_cannot_be_parsed + stop (:161-172)
        die("ERROR: Not enough coefficients in the SC wavefunction file to unfold pc-kpt #" +
            std::to_string(rows[(size_t)f] + 1) + " (it cannot be parsed).");
      }
      std::memcpy(&dN_all[(size_t)rows[(size_t)f] * (size_t)nener], &dN[(size_t)f * (size_t)nener], sizeof(double) * (size_t)nener);
      ++n_parsed;
    }
    const bool this_spinor = spinor;
    if (write_udo || this_spinor) {
      int64_t n_e = 0, n_x = 0;
      gpu_check(h, bandup_gpu_rho_count(h, j + 1, 1e-3, &n_e, &n_x), "calc_rho");
      if (write_udo && n_x > 0) {
        std::vector<int32_t> fo((size_t)n_x), ie((size_t)n_x), m1((size_t)n_x), m2((size_t)n_x);
        std::vector<float> re((size_t)n_x), im((size_t)n_x);
        gpu_check(h, bandup_gpu_fetch_rho(h, j + 1, n_x, fo.data(), ie.data(), m1.data(), m2.data(), re.data(), im.data()), "fetch_rho");
        for (int64_t x = 0; x < n_x; ++x) {
          RhoRows& r = rho_of[(size_t)rows[(size_t)fo[(size_t)x]]];
          r.ie.push_back(ie[(size_t)x]); r.m1.push_back(m1[(size_t)x]); r.m2.push_back(m2[(size_t)x]);
          r.re.push_back(re[(size_t)x]); r.im.push_back(im[(size_t)x]);
        }
      }
      if (this_spinor) {
        if (sigma_all.empty()) sigma_all.assign(n_pck * (size_t)nener * 3, 0.0);
        std::vector<double> sg((size_t)nf * (size_t)nener * 3);
        gpu_check(h, bandup_gpu_pauli_vectors(h, j + 1, sg.data()), "pauli_vectors");
        for (int f = 0; f < nf; ++f)
          std::memcpy(&sigma_all[(size_t)rows[(size_t)f] * (size_t)nener * 3], &sg[(size_t)f * (size_t)nener * 3],
                      sizeof(double) * (size_t)nener * 3);
      }
    }
  };

  std::future<void> ahead;
  if (next_used())
    ahead = std::async(std::launch::async, [&] {
      const double t0 = now_s();
      wc.read_band_records(used[0] + 1, spin_channel, bufs[0], read_threads);
      t_read += now_s() - t0;
    });
  for (size_t j = 0; j < used.size(); ++j) {
    ahead.get();
    if (next_used()) {  // one k-point ahead: its records are read while k-point j uploads / computes
      const int i_next = used[j + 1];
      ahead = std::async(std::launch::async, [&, j, i_next] {
        const double t0 = now_s();
        wc.read_band_records(i_next + 1, spin_channel, bufs[(j + 1) % (size_t)n_buf], read_threads);
        t_read += now_s() - t0;
      });
    }
    if ((int)inflight.size() >= depth) {
      drain(inflight.front());
      inflight.erase(inflight.begin());
    }
    const int i = used[j];
    const auto& rows = folds[(size_t)i];
    const int nf = (int)rows.size();
    std::vector<double> fG((size_t)nf * 3);  // folding_G = k - K (band_unfolding_routines_mod.f90:571-576)
    for (int f = 0; f < nf; ++f)
      for (int c = 0; c < 3; ++c) fG[(size_t)f * 3 + (size_t)c] = pck[(size_t)rows[(size_t)f]][(size_t)c] - Kcart[(size_t)i][(size_t)c];
    std::vector<int32_t> g0((size_t)nf * 3);
    double resid = 0.0;
    gpu_check(h, bandup_gpu_folding_g0(h, fG.data(), nf, g0.data(), &resid), "select_coeffs_to_calc_spectral_weights");
    int nsp = 1, npw = 0;
    gpu_check(h, bandup_gpu_submit_kpt_vasp(h, (int)j + 1, nplane[(size_t)i], wc.nbands, wc.nrecl, bufs[j % (size_t)n_buf],
                                            Kfrac[(size_t)i].data(), wc.encut, 0.0, ener[(size_t)i].data(), nf, g0.data(),
                                            &nsp, &npw), "read_wavefunction");
    spinor = spinor || nsp == 2;
    inflight.push_back((int)j);
    std::printf("SC-kpt #%d: %d plane waves%s, %d pc-kpt(s) fold onto it\n", i + 1, npw, nsp == 2 ? " (spinor)" : "", nf);
  }
  for (int j : inflight) drain(j);
  size_t n_pck_mine = 0;
  for (int i : used) n_pck_mine += folds[(size_t)i].size();
  if (n_parsed != n_pck_mine) die("ERROR: not every pc-kpt could be unfolded.");
  if (n_ranks > 1) {  // the path's only collective: every rank's disjoint delta_N rows (ncclAllGather)
    std::vector<double> mine(n_pck_mine * (size_t)nener), all(n_pck * (size_t)nener);
    std::vector<int64_t> ids, all_ids(n_pck);
    for (int i : used)
      for (int r : folds[(size_t)i]) {
        std::memcpy(&mine[ids.size() * (size_t)nener], &dN_all[(size_t)r * (size_t)nener], sizeof(double) * (size_t)nener);
        ids.push_back(r);
      }
    int64_t n_total = 0;
    gpu_check(h, bandup_gpu_gather_dN(h, mine.data(), ids.data(), (int64_t)ids.size(), all.data(), all_ids.data(),
                                      (int64_t)n_pck, &n_total), "gather_dN");
    if ((size_t)n_total != n_pck) die("ERROR: not every pc-kpt could be unfolded.");
    for (size_t r = 0; r < n_pck; ++r)
      std::memcpy(&dN_all[(size_t)all_ids[r] * (size_t)nener], &all[r * (size_t)nener], sizeof(double) * (size_t)nener);
    // the unfolding-density operators have no fixed row length: the other ranks leave theirs in a
    // part file next to the output (one node, one file system) and rank 0 merges them after the
    // next collective, which doubles as the barrier
    const auto part_name = [&](int r) { return udo_file + ".part" + std::to_string(r); };
    if (write_udo && rank != 0) {
      std::FILE* pf = std::fopen(part_name(rank).c_str(), "wb");
      if (!pf) die("ERROR: cannot write " + part_name(rank));
      for (int64_t r : ids) {
        const RhoRows& rr = rho_of[(size_t)r];
        const int64_t head[2] = {r, (int64_t)rr.ie.size()};
        std::fwrite(head, sizeof(int64_t), 2, pf);
        std::fwrite(rr.ie.data(), sizeof(int32_t), rr.ie.size(), pf);
        std::fwrite(rr.m1.data(), sizeof(int32_t), rr.m1.size(), pf);
        std::fwrite(rr.m2.data(), sizeof(int32_t), rr.m2.size(), pf);
        std::fwrite(rr.re.data(), sizeof(float), rr.re.size(), pf);
        std::fwrite(rr.im.data(), sizeof(float), rr.im.size(), pf);
      }
      if (std::fclose(pf) != 0) die("ERROR: cannot write " + part_name(rank));
    }
    {  // spinor or not is learnt from the k-points a rank submitted: agree on it (a rank may have none)
      double flag = spinor ? 1.0 : 0.0, flags[1024];
      int64_t me = rank, who[1024], n_flags = 0;
      gpu_check(h, bandup_gpu_gather_rows(h, 1, &flag, &me, 1, flags, who, 1024, &n_flags), "gather_rows");
      for (int64_t r = 0; r < n_flags; ++r) spinor = spinor || flags[r] != 0.0;
    }
    if (write_udo && rank == 0) {
      for (int r = 1; r < n_ranks; ++r) {
        std::FILE* pf = std::fopen(part_name(r).c_str(), "rb");
        if (!pf) die("ERROR: cannot read " + part_name(r));
        int64_t head[2];
        while (std::fread(head, sizeof(int64_t), 2, pf) == 2) {
          if (head[0] < 0 || (size_t)head[0] >= n_pck || head[1] < 0) die("ERROR: corrupt " + part_name(r));
          RhoRows& rr = rho_of[(size_t)head[0]];
          const size_t n = (size_t)head[1];
          rr.ie.resize(n); rr.m1.resize(n); rr.m2.resize(n); rr.re.resize(n); rr.im.resize(n);
          if (std::fread(rr.ie.data(), sizeof(int32_t), n, pf) != n || std::fread(rr.m1.data(), sizeof(int32_t), n, pf) != n ||
              std::fread(rr.m2.data(), sizeof(int32_t), n, pf) != n || std::fread(rr.re.data(), sizeof(float), n, pf) != n ||
              std::fread(rr.im.data(), sizeof(float), n, pf) != n)
            die("ERROR: truncated " + part_name(r));
        }
        std::fclose(pf);
        std::remove(part_name(r).c_str());
      }
    }
    if (spinor) {  // the spin columns travel the same way: rows of 3*nener doubles
      const size_t rl = (size_t)nener * 3;
      if (sigma_all.empty()) sigma_all.assign(n_pck * rl, 0.0);
      std::vector<double> smine(n_pck_mine * rl), sall(n_pck * rl);
      for (size_t k = 0; k < ids.size(); ++k)
        std::memcpy(&smine[k * rl], &sigma_all[(size_t)ids[k] * rl], sizeof(double) * rl);
      gpu_check(h, bandup_gpu_gather_rows(h, (int)rl, smine.data(), ids.data(), (int64_t)ids.size(), sall.data(),
                                          all_ids.data(), (int64_t)n_pck, &n_total), "gather_rows");
      for (size_t r = 0; r < n_pck; ++r)
        std::memcpy(&sigma_all[(size_t)all_ids[r] * rl], &sall[r * rl], sizeof(double) * rl);
    }
    gpu_check(h, bandup_gpu_comm_finalize(h), "comm_finalize");
    if (rank == 0) {  // every rank is past its last ticket and the collectives: the rendezvous files can go
      std::remove(args.get("-nccl_id_file", "bandup_gpu_nccl_id").c_str());
      if (!ticket_file.empty()) std::remove(ticket_file.c_str());
    }
    if (rank != 0) {  // rank 0 writes the files
      for (auto b : bufs) bandup_gpu_pinned_free(b);
      bandup_gpu_finalize(h);
      return 0;
    }
  }

  // write_band_struc (io_routines_mod.f90:889-977): '(2(f8.4,2X),ES10.3)' [+ 5 spin columns]
  const double t_write0 = now_s();
  double normal[3] = {0, 0, 1}, origin[3] = {0, 0, 0};
  parse_vec3(args.get("-normal_to_proj_plane", "0, 0, 1"), normal);
  parse_vec3(args.get("-origin_for_spin_proj_cart", "0, 0, 0"), origin);
  if (args.has_value("-origin_for_spin_proj_rec")) {
    // fractional w.r.t. the SC reciprocal lattice; it wins over the Cartesian one as in the
    // reference (cla_wrappers_mod.f90:452-467, band_unfolding_routines_mod.f90:183-189)
    double rec[3];
    parse_vec3(args.get("-origin_for_spin_proj_rec", "0, 0, 0"), rec);
    vec_mat(rec, B_sc, origin);
  }
  {
    std::FILE* f = std::fopen(out_file.c_str(), "w");
    if (!f) die("ERROR: cannot write " + out_file);
    std::fprintf(f, "# File created by BandUP_gpu.x (bandup-b200, GPU unfolding path)\n");
    if (spinor)
    {  // the three comment lines of the reference's spinor output (io_routines_mod.f90:921-929)
      const double sax[3] = {0, 0, 1};  // the reference resets args%saxis to (0,0,1) whatever -saxis says (cla_wrappers_mod.f90:436)
      std::fprintf(f, "# NB.: Unfolding of Pauli matrices eigenvs is a test feature.\n");
      std::fprintf(f, "# The quantization axis assumed was saxis = [%s, %s, %s]\n", es10_3(sax[0]).c_str(),
                   es10_3(sax[1]).c_str(), es10_3(sax[2]).c_str());
      std::fprintf(f, "# The projection axis used for \"#Spin _|_ k\" is also _|_ to [%s, %s, %s].\n",
                   es10_3(normal[0]).c_str(), es10_3(normal[1]).c_str(), es10_3(normal[2]).c_str());
      // '(3(A,1X), 2X, 5(A,1X))' (:930-933): the last label carries its own blank and is followed by 1X
      std::fprintf(f, "#KptCoord #E-E_Fermi #delta_N   #sigma_x    #sigma_y    #sigma_z    #Spin _|_ k #Spin // k  \n");
    }
    else
      std::fprintf(f, "#KptCoord #E-E_Fermi #delta_N \n");
    double coord_first = kp.zero, coord_k = kp.zero;
    size_t row = 0;
    for (auto& d : kp.dirs) {
      for (auto& k : d) {
        coord_k = coord_first + std::sqrt((k[0] - d[0][0]) * (k[0] - d[0][0]) + (k[1] - d[0][1]) * (k[1] - d[0][1]) +
                                          (k[2] - d[0][2]) * (k[2] - d[0][2]));
        for (int ie = 0; ie < nener; ++ie) {
          std::fprintf(f, "%8.4f  %8.4f  %s", coord_k, grid[(size_t)ie] - ef, es10_3(dN_all[row * (size_t)nener + (size_t)ie]).c_str());
          if (spinor) {
            const double* s = &sigma_all[(row * (size_t)nener + (size_t)ie) * 3];
            double perp = 0.0, para = 0.0;  // calc_spin_projections (band_unfolding_routines_mod.f90:1266-1286)
            double kv[3] = {k[0] - origin[0], k[1] - origin[1], k[2] - origin[2]};
            const double kn = std::sqrt(kv[0] * kv[0] + kv[1] * kv[1] + kv[2] * kv[2]);
            const double nn = std::sqrt(normal[0] * normal[0] + normal[1] * normal[1] + normal[2] * normal[2]);
            if (kn > 1e-12 && nn > 0.0 && (s[0] != 0.0 || s[1] != 0.0 || s[2] != 0.0)) {
              double vpara[3] = {kv[0] / kn, kv[1] / kn, kv[2] / kn}, nv[3] = {normal[0] / nn, normal[1] / nn, normal[2] / nn}, vperp[3];
              cross(nv, vpara, vperp);
              const double pn = std::sqrt(vperp[0] * vperp[0] + vperp[1] * vperp[1] + vperp[2] * vperp[2]);
              if (pn > 0.0) perp = (s[0] * vperp[0] + s[1] * vperp[1] + s[2] * vperp[2]) / pn;
              para = s[0] * vpara[0] + s[1] * vpara[1] + s[2] * vpara[2];
            }
            std::fprintf(f, "  %s  %s  %s  %s  %s  ", es10_3(s[0]).c_str(), es10_3(s[1]).c_str(), es10_3(s[2]).c_str(),
                         es10_3(perp).c_str(), es10_3(para).c_str());
          }
          std::fprintf(f, "\n");
        }
        ++row;
      }
      coord_first = coord_k;
    }
    std::fclose(f);
  }

  // write_unf_dens_op (io_routines_mod.f90:1637-1882): upper triangle, |rho| >= 1e-4, next to rho*N
  if (write_udo) {
    std::FILE* f = std::fopen(udo_file.c_str(), "w");
    if (!f) die("ERROR: cannot write " + udo_file);
    const std::string bar(88, '#');
    std::fprintf(f, "%s\n# File created by BandUP_gpu.x (bandup-b200, GPU unfolding path)\n%s\n", bar.c_str(), bar.c_str());
    std::fprintf(f, "# Both UnfDensOp and GenSpecWeight are hermitian; only the upper-triangular matrix\n"
                    "# elements are present. They are represented as (RealPart, ImagPart), and were\n"
                    "# evaluated between SC states with band indices indicated by m1 and m2.\n#\n");
    std::fprintf(f, "# SpinChannel = %d\n# nScBands = %d\n# emin = %.5f eV\n# emax = %.5f eV\n# nEner = %d\n"
                    "# OriginalEfermi = %.5f eV\n# The energies were shifted so that the Fermi level is at 0 eV.\n%s\n",
                 spin_channel, wc.nbands, grid.front() - ef, grid.back() - ef, nener, ef, bar.c_str());
    Mat3 binv_pc;
    inverse(b_pc, &binv_pc);
    double coord_first = 0.0, coord_k = 0.0;
    size_t row = 0;
    for (auto& d : kp.dirs) {
      for (auto& k : d) {
        coord_k = coord_first + std::sqrt((k[0] - d[0][0]) * (k[0] - d[0][0]) + (k[1] - d[0][1]) * (k[1] - d[0][1]) +
                                          (k[2] - d[0][2]) * (k[2] - d[0][2]));
        const int isc = sck_of[row];
        double fs[3], fp[3], Kf[3];
        vec_mat(k.data(), Binv, fs);
        vec_mat(k.data(), binv_pc, fp);
        vec_mat(Kcart[(size_t)isc].data(), Binv, Kf);
        std::fprintf(f, "\n# PcKptNumber = %zu\n", row + 1);
        std::fprintf(f, "#     PcKptCartesianCoords =  %.8f  %.8f  %.8f  (A^{-1})\n", k[0], k[1], k[2]);
        std::fprintf(f, "#     PcKptFractionalCoords =  %.8f  %.8f  %.8f  (w.r.t. SCRL basis)\n", fs[0], fs[1], fs[2]);
        std::fprintf(f, "#     PcKptFractionalCoords =  %.8f  %.8f  %.8f  (w.r.t. PCRL basis)\n", fp[0], fp[1], fp[2]);
        std::fprintf(f, "#     PcKptLinearCoordsInBandPlot = %.4f  (A^{-1})\n", coord_k);
        std::fprintf(f, "#     Folds into ScKptNumber = %d\n", isc + 1);
        std::fprintf(f, "#         ScKptCartesianCoords =  %.8f  %.8f  %.8f  (A^{-1})\n", Kcart[(size_t)isc][0], Kcart[(size_t)isc][1], Kcart[(size_t)isc][2]);
        std::fprintf(f, "#         ScKptFractionalCoords =  %.8f  %.8f  %.8f  (w.r.t. SCRL basis)\n#\n", Kf[0], Kf[1], Kf[2]);
        const RhoRows& r = rho_of[row];
        size_t x = 0;
        while (x < r.ie.size()) {  // entries arrive grouped by energy, m1 outer / m2 inner
          size_t y = x;
          while (y < r.ie.size() && r.ie[y] == r.ie[x]) ++y;
          const int ie = r.ie[x];
          const double N = dN_all[row * (size_t)nener + (size_t)(ie - 1)];
          if (N >= 1e-3) {
            double trace = 0.0;
            for (size_t z = x; z < y; ++z)
              if (r.m1[z] == r.m2[z]) trace += r.re[z];
            char tr[32];
            std::snprintf(tr, sizeof tr, "%8.1E", trace);
            std::string trs = tr;
            trs.erase(0, trs.find_first_not_of(' '));
            if (!(std::fabs(trace - 1.0) < 1e-1)) trs += " (!=1)";
            std::fprintf(f, "#     iEnerPC = %d  E = %.4f eV  N(k,E) = %s  Tr{UnfDensOp} = %s\n", ie, grid[(size_t)(ie - 1)], es10_3(N).c_str(), trs.c_str());
            std::fprintf(f, "#        m1      m2        GenSpecWeight[m1,m2]         UnfDensOp[m1,m2]\n");
            for (size_t z = x; z < y; ++z) {
              if (r.m2[z] < r.m1[z] || std::hypot(r.re[z], r.im[z]) < 1e-4f) continue;
              std::fprintf(f, "      %5d   %5d      (%s, %s)   (%s, %s)\n", r.m1[z], r.m2[z], es10_3((double)r.re[z] * N).c_str(),
                           es10_3((double)r.im[z] * N).c_str(), es10_3(r.re[z]).c_str(), es10_3(r.im[z]).c_str());
            }
          }
          x = y;
        }
        ++row;
      }
      coord_first = coord_k;
    }
    std::fclose(f);
  }

  t_write = now_s() - t_write0;
  std::printf("%zu pc-kpts unfolded from %zu SC-kpts; %lld GPU kernel launches.\n", n_pck, used.size(),
              (long long)bandup_gpu_launch_count(h));
  {  // print_final_times (io_routines_mod.f90:1923-1962), the categories this host has
    double ms[BANDUP_GPU_K_COUNT] = {0};
    for (int k = 0; k < BANDUP_GPU_K_COUNT; ++k) {
      int64_t n = 0;
      bandup_gpu_kernel_time(h, k, &ms[k], &n);
    }
    const double total = now_s() - t_start;
    print_time("Total elapsed time:                                               ", total, total);
    print_time("Time spent reading the wavefunctions file (overlapped):           ", t_read, total);
    print_time("Time spent calculating spectral weights (GPU kernels):            ",
               1e-3 * (ms[BANDUP_GPU_K_COSET] + ms[BANDUP_GPU_K_WEIGHTS] + ms[BANDUP_GPU_K_FINALIZE]), total);
    print_time("Time spent calculating the unfolded N(k,E) (GPU kernels):         ", 1e-3 * ms[BANDUP_GPU_K_DELTA_N], total);
    print_time("Time spent calculating unfolding-density operators (GPU kernels): ", 1e-3 * ms[BANDUP_GPU_K_RHO], total);
    print_time("Time spent writing output file(s):                                ", t_write, total);
  }
  for (auto b : bufs) bandup_gpu_pinned_free(b);
  bandup_gpu_finalize(h);
  ::close(wc.fd);
  return 0;
}
