#include "cafe_commands.h"

#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
#include <sstream>
#include <stdexcept>

Globals::Globals() {
    param.family_size = family_size_range{0, 1, 0, 1};
    param.optimizer_init_type = LAMBDA_ONLY;
}

Globals::~Globals() { Clear(0); }

void Globals::Clear(int btree_skip) {
    if (param.pfamily) { cafe_family_free(param.pfamily); param.pfamily = nullptr; }
    if (!btree_skip && param.pcafe) { cafe_tree_free(param.pcafe); param.pcafe = nullptr; }
    if (param.flog && param.flog != stdout && param.flog != stderr) { fclose(param.flog); }
    param.flog = stdout;
    param.cond_dist.clear();
    param.max_pvalues.clear();
}

void Globals::Prepare() {
    param.lambda = nullptr;
    param.mu = nullptr;
    param.lambda_tree.clear();
    param.lambda_tree_string.clear();
    param.num_lambdas = -1;
    param.num_mus = -1;
    param.parameterized_k_value = 0;
    param.optimizer_init_type = LAMBDA_ONLY;
}

std::vector<std::string> tokenize(const std::string& s) {
    std::istringstream iss(s);
    std::vector<std::string> out;
    std::string tok;
    while (iss >> tok) out.push_back(tok);
    return out;
}

std::vector<Argument> build_argument_list(std::vector<std::string> tokens) {
    auto is_opt = [](const std::string& t) { return t.size() > 1 && t[0] == '-' && !std::isdigit((unsigned char)t[1]); };
    std::vector<Argument> result;
    for (size_t i = 1; i < tokens.size(); ++i) {
        if (!is_opt(tokens[i])) continue;
        Argument a;
        a.opt = tokens[i];
        size_t j = i + 1;
        for (; j < tokens.size() && !is_opt(tokens[j]); ++j) a.argv.push_back(tokens[j]);
        result.push_back(a);
        i = j - 1;
    }
    return result;
}

static void prereqs(pCafeParam param, bool family, bool tree, bool lambda) {
    if (family && !param->pfamily)
        throw std::runtime_error("ERROR: The gene families were not loaded. Please load gene families with the 'load' command.\n");
    if (tree && !param->pcafe)
        throw std::runtime_error("ERROR: The tree was not loaded. Please load a tree with the 'tree' command.\n");
    if (lambda && !param->lambda)
        throw std::runtime_error("ERROR: Lambda values were not set. Please set lambda values with the 'lambda' or 'lambdamu' commands.\n");
}

static std::string join(const std::vector<std::string>& v, const char* sep) {
    std::string s;
    for (size_t i = 0; i < v.size(); ++i) { if (i) s += sep; s += v[i]; }
    return s;
}

int cafe_cmd_seed(Globals&, std::vector<std::string> tokens) {
    if (tokens.size() < 2) throw std::runtime_error("No value provided for seed");
    try { std::srand(std::stoi(tokens[1])); }
    catch (...) { throw std::runtime_error("Failed to set seed from value " + tokens[1]); }
    return 0;
}

int cafe_cmd_load(Globals& globals, std::vector<std::string> tokens) {
    pCafeParam param = &globals.param;
    if (tokens.size() < 2) throw std::runtime_error("Usage(load): load <family file>\n");
    globals.Clear(1);
    std::string file, logfile;
    int max_size = -1;
    for (const Argument& a : build_argument_list(tokens)) {
        if (a.opt == "-t" && a.argc()) { int t = 0; std::sscanf(a.argv[0].c_str(), "%d", &t); if (t > 0) param->num_threads = t; }
        if (a.opt == "-r" && a.argc()) { int r = 0; std::sscanf(a.argv[0].c_str(), "%d", &r); if (r > 0) globals.num_random_samples = r; }
        if (a.opt == "-max_size" && a.argc()) std::sscanf(a.argv[0].c_str(), "%d", &max_size);
        if (a.opt == "-p" && a.argc()) { double p = -1; std::sscanf(a.argv[0].c_str(), "%lf", &p); if (p > 0) param->pvalue = p; }
        if (a.opt == "-l" && a.argc()) logfile = join(a.argv, " ");
        if (a.opt == "-i") file = join(a.argv, " ");
        if (a.opt == "-filter") std::cerr << "load -filter is outside the GPU hot path; ignored\n";
    }
    param->num_random_samples = globals.num_random_samples;
    if (!logfile.empty() && logfile != "stdout") {
        param->flog = std::fopen(logfile.c_str(), "a");
        if (!param->flog) { param->flog = stdout; throw std::runtime_error("ERROR(load): Cannot open log file: " + logfile); }
    }
    if (file.empty()) throw std::runtime_error("ERROR(load): You must use -i option for input file\n");
    char sep = (file.size() >= 3 && file.compare(file.size() - 3, 3, "csv") == 0) ? ',' : '\t';
    param->str_fdata = file;
    std::ifstream ifst(file);
    param->pfamily = load_gene_families(ifst, sep, max_size);
    if (!param->pfamily) throw std::runtime_error("Failed to load file\n");
    init_family_size(&param->family_size, param->pfamily->max_size);
    if (param->pcafe) {
        cafe_tree_set_parameters(param->pcafe, &param->family_size, 0);
        cafe_family_set_species_index(param->pfamily, param->pcafe);
    }
    return 0;
}

int cafe_cmd_tree(Globals& globals, std::vector<std::string> tokens) {
    pCafeParam param = &globals.param;
    std::string newick;
    if (tokens.size() == 1) throw std::runtime_error("Failed to read input\n");
    if (tokens.size() > 2 && tokens[1] == "-i") {
        std::ifstream ifst(tokens[2]);
        if (!ifst) throw std::runtime_error("Failed to read file '" + tokens[2] + "'");
        std::stringstream buffer;
        buffer << ifst.rdbuf();
        newick = buffer.str();
    } else {
        for (size_t i = 1; i < tokens.size(); ++i) newick += tokens[i];
    }
    if (param->pcafe) { cafe_free_birthdeath_cache(param->pcafe); cafe_tree_free(param->pcafe); param->pcafe = nullptr; }
    param->pcafe = cafe_tree_new(newick.c_str(), &param->family_size, 0, 0);
    if (!param->pcafe) throw std::runtime_error("Failed to load tree from provided string");
    if (!is_ultrametric(param->pcafe)) std::cerr << "WARNING: tree is not ultrametric\n";
    param->num_branches = param->pcafe->num_nodes() - 1;
    max_branch_length(param->pcafe);  // throws when a branch length is missing
    if (!param->quiet) std::printf("%s\n", cafe_tree_string(*param->pcafe).c_str());
    if (param->pfamily) cafe_family_set_species_index(param->pfamily, param->pcafe);
    return 0;
}

static void set_prior(Globals& globals) {
    std::vector<double> prior;
    cafe_set_prior_rfsize_empirical(&globals.param, prior);
    globals.param.prior_rfsize_store = prior;
    globals.param.prior_rfsize = globals.param.prior_rfsize_store.data();
}

static std::vector<double> doubles_of(const Argument& a) {
    std::vector<double> v;
    for (const auto& s : a.argv) { double d = 0; std::sscanf(s.c_str(), "%lf", &d); v.push_back(d); }
    return v;
}

int cafe_cmd_lambda(Globals& globals, std::vector<std::string> tokens) {
    pCafeParam param = &globals.param;
    prereqs(param, true, true, false);
    std::vector<Argument> pargs = build_argument_list(tokens);
    globals.Prepare();
    bool search = false, score = false, have_tree = false, checkconv = false;
    std::vector<double> lambdas;
    int given = 0;
    for (const Argument& a : pargs) {
        if (a.opt == "-s") search = true;
        else if (a.opt == "-checkconv") checkconv = true;
        else if (a.opt == "-score") score = true;
        else if (a.opt == "-t") {
            if (a.argc() < 1) throw std::runtime_error("lambda -t needs a tree");
            if (__cafe_cmd_lambda_tree(param, a.argv[0].c_str(), a.argc() > 1 ? a.argv[1].c_str() : nullptr) < 0) throw std::exception();
            cafe_log(param, "Lambda Tree: %s\n", param->lambda_tree_string.c_str());
            have_tree = true;
            lambdas.resize(param->num_lambdas);
        } else if (a.opt == "-l") {
            lambdas = doubles_of(a);
            given += (int)lambdas.size();
        } else if (a.opt == "-k" || a.opt == "-p" || a.opt == "-f") {
            throw std::runtime_error("lambda -k/-p/-f (clustered model) is outside the GPU hot path");
        } else if (a.opt == "-r" || a.opt == "-e" || a.opt == "-o" || a.opt == "-v") {
            throw std::runtime_error("lambda " + a.opt + " is outside the GPU hot path");
        }
    }
    param->posterior = 1;
    set_prior(globals);
    if (search) {  // lambda_search, lambda.cpp:324-355
        if (!have_tree) { param->num_lambdas = 1; lambdas.resize(1); }
        param->num_params = param->num_lambdas;
        param->input.construct(param->num_params);
        if (checkconv) param->checkconv = 1;
        cafe_best_lambda_by_fminsearch(param, param->num_lambdas, 0);
    } else {  // lambda_set, lambda.cpp:285-322
        if (!have_tree) param->num_lambdas = 1;
        param->num_params = param->num_lambdas;
        if (given != param->num_params) {
            std::ostringstream ost;
            ost << "ERROR (lambda): The total number of parameters was not correct.\nThe total number of lambdas (-l) are "
                << given << " but " << param->num_params << " were expected\n";
            throw std::runtime_error(ost.str());
        }
        param->input.construct(param->num_params);
        for (int i = 0; i < param->num_params; ++i) param->input.parameters[i] = lambdas[i];
        cafe_shell_set_lambda(param, param->input.parameters);
        if (score) __cafe_best_lambda_search(param->lambda, param);
    }
    if (param->pfamily) reset_birthdeath_cache(param->pcafe, param->parameterized_k_value, &param->family_size);
    cafe_log(param, "DONE: Lambda Search or setting, for command:\n");
    std::ostringstream cmd;
    std::copy(tokens.begin(), tokens.end(), std::ostream_iterator<std::string>(cmd, " "));
    cafe_log(param, "%s\n", cmd.str().c_str());
    return 0;
}

int cafe_cmd_lambdamu(Globals& globals, std::vector<std::string> tokens) {
    pCafeParam param = &globals.param;
    prereqs(param, true, true, false);
    std::vector<Argument> pargs = build_argument_list(tokens);
    globals.Prepare();
    param->optimizer_init_type = LAMBDA_MU;
    bool search = false, have_tree = false, checkconv = false;
    int eqbg = 0, given = 0;
    std::vector<double> lambdas, mus;
    for (const Argument& a : pargs) {
        if (a.opt == "-s") search = true;
        else if (a.opt == "-checkconv") checkconv = true;
        else if (a.opt == "-t") {
            if (a.argc() < 1) throw std::runtime_error("lambdamu -t needs a tree");
            if (__cafe_cmd_lambda_tree(param, a.argv[0].c_str(), a.argc() > 1 ? a.argv[1].c_str() : nullptr) < 0) throw std::exception();
            cafe_log(param, "Lambda Tree: %s\n", param->lambda_tree_string.c_str());
            have_tree = true;
            lambdas.resize(param->num_lambdas);
            param->num_mus = param->num_lambdas;
            mus.resize(param->num_mus);
        } else if (a.opt == "-l") { lambdas = doubles_of(a); given += (int)lambdas.size(); }
        else if (a.opt == "-m") { mus = doubles_of(a); given += (int)mus.size(); }
        else if (a.opt == "-eqbg") eqbg = 1;
        else if (a.opt == "-k" || a.opt == "-p" || a.opt == "-f") throw std::runtime_error("lambdamu -k/-p/-f (clustered model) is outside the GPU hot path");
    }
    param->posterior = 1;
    set_prior(globals);
    std::ostream& log = std::cout;
    if (have_tree) {
        param->eqbg = eqbg;
    } else {
        param->num_lambdas = 1; lambdas.resize(1);
        param->num_mus = 1; mus.resize(1);
        if (eqbg) throw std::runtime_error("ERROR(lambdamu): Cannot use option eqbg without specifying a lambda tree. \n");
    }
    // lambdamu_args::get_num_params, lambdamu.cpp:48-75 (k = 0)
    const int expected = have_tree ? (int)(lambdas.size() + mus.size()) - (search ? 0 : 0) : (int)lambdas.size() + ((int)mus.size() - eqbg);
    if (search) {
        param->num_params = have_tree ? (int)(lambdas.size() + mus.size()) : (int)lambdas.size() + ((int)mus.size() - eqbg);
        param->input.construct(param->num_params);
        if (checkconv) param->checkconv = 1;
        best_lambda_mu_by_fminsearch(param, param->num_lambdas, param->num_mus, 0, log);
    } else {
        param->num_params = have_tree ? (int)lambdas.size() + ((int)mus.size() - eqbg) : (int)(lambdas.size() + mus.size());
        (void)expected;
        if (given != param->num_params) {
            std::ostringstream ost;
            ost << "ERROR (lambdamu): The total number of parameters was not correct.\nThe total number of lambdas (-l) and mus (-m) are "
                << given << " but " << param->num_params << " were expected\n";
            throw std::runtime_error(ost.str());
        }
        param->input.construct(param->num_params);
        for (size_t i = 0; i < lambdas.size(); ++i) param->input.parameters[i] = lambdas[i];
        for (int i = 0; i < (int)mus.size() - (have_tree ? eqbg : 0); ++i) param->input.parameters[param->num_lambdas + i] = mus[i];
        cafe_shell_set_lambda_mu(param, param->input.parameters);
    }
    if (param->pfamily) reset_birthdeath_cache(param->pcafe, param->parameterized_k_value, &param->family_size);
    cafe_log(param, "DONE: Lamda,Mu Search or setting, for command:\n");
    std::ostringstream ost;
    std::copy(tokens.begin(), tokens.end(), std::ostream_iterator<std::string>(ost, " "));
    cafe_log(param, "%s\n", ost.str().c_str());
    return 0;
}

int cafe_cmd_errormodel(Globals& globals, std::vector<std::string> tokens) {
    pCafeParam param = &globals.param;
    prereqs(param, true, true, false);
    std::string model_file;
    std::vector<std::string> species;
    bool all = false;
    for (const Argument& a : build_argument_list(tokens)) {
        if (a.opt == "-model" && a.argc()) model_file = a.argv[0];
        if (a.opt == "-sp") species = a.argv;
        if (a.opt == "-all") all = true;
    }
    if (!model_file.empty()) {
        if (!species.empty()) {
            for (const auto& sp : species) set_error_matrix_from_file(param->pfamily, param->pcafe, param->family_size, model_file, sp);
        } else if (all) {
            set_error_matrix_from_file(param->pfamily, param->pcafe, param->family_size, model_file, std::string());
        }
        std::fprintf(stderr, "errormodel: %s set.\n", model_file.c_str());
    }
    if (param->pfamily->errors.empty())
        throw std::runtime_error("ERROR(errormodel): we need an error model specified (-model) or two data files.\n");
    return 0;
}

int cafe_cmd_pvalue(Globals& globals, std::vector<std::string> tokens) {
    pCafeParam param = &globals.param;
    std::string outfile, infile;
    for (const Argument& a : build_argument_list(tokens)) {
        if (a.opt == "-o" && a.argc()) outfile = a.argv[0];
        if (a.opt == "-i" && a.argc()) infile = a.argv[0];
    }
    if (!outfile.empty()) {
        prereqs(param, false, true, true);
        std::ofstream ofst(outfile.c_str());
        if (!ofst) throw std::runtime_error("ERROR(pvalue): Cannot open " + outfile + " in write mode.\n");
        param->cond_dist = cafe_conditional_distribution(param->pcafe, &param->family_size, param->num_threads, globals.num_random_samples);
        write_pvalues(ofst, param->cond_dist, globals.num_random_samples);
    } else if (!infile.empty()) {
        cafe_log(param, "Loading p-values ... \n");
        std::ifstream ifst(infile.c_str());
        if (!ifst) throw std::runtime_error("ERROR(pvalue): Cannot open " + infile + " in read mode.\n");
        param->cond_dist = read_pvalues(ifst, globals.num_random_samples);
        cafe_log(param, "Done Loading p-values ... \n");
    } else {
        throw std::runtime_error("pvalue: only -o <file> / -i <file> are on the GPU hot path");
    }
    return 0;
}

// cafe_do_report (reports.cpp:650-708): matrices at the current lambda, conditional distribution, then cafe_viterbi (family-wide
// p-values, Viterbi reconstruction, branch p-values), optionally the likelihood-ratio test, and the text report "<name>.cafe" in
// the reference's format (cafe_report_text).  "<name>.pvalues" (one "ID<TAB>p" line per family) is kept as a convenience.
// Options: `likelihood` = the test with the nodes' own mu; `likelihood-stock` = keyed like the stock binary (cafe_param.h);
// `branchcutting` = the branch-cutting p-values (the nodes' own mu; `branchcutting-stock` keys the copies like the stock binary would);
// `lh2` fails inside the reference itself and is rejected; html/json formats are not built.
int cafe_cmd_report(Globals& globals, std::vector<std::string> tokens) {
    pCafeParam param = &globals.param;
    prereqs(param, true, true, true);
    if (tokens.size() < 2) throw std::runtime_error("Usage(report): report <name>\n");
    bool likelihood = false, branchcutting = false;
    for (size_t i = 2; i < tokens.size(); ++i) {  // report_parameters, reports.cpp:604-616
        std::string o = tokens[i];
        for (char& c : o) c = (char)std::tolower((unsigned char)c);
        if (o == "likelihood") { likelihood = true; param->lrt_tree_level_mu = 0; }
        if (o == "likelihood-stock") { likelihood = true; param->lrt_tree_level_mu = 1; }
        if (o == "branchcutting") { branchcutting = true; param->lrt_tree_level_mu = 0; }
        if (o == "branchcutting-stock") { branchcutting = true; param->lrt_tree_level_mu = 1; }
        if (o == "lh2" || o == "html" || o == "json") throw std::runtime_error("report: " + o + " is not built (SURVEY.md 8f)");
    }
    cafe_shell_set_lambdas(param, param->input.parameters);
    reset_birthdeath_cache(param->pcafe, param->parameterized_k_value, &param->family_size);
    if (param->cond_dist.empty()) {
        cafe_log(param, "Running Conditional Distribution ...\n");
        param->cond_dist = cafe_conditional_distribution(param->pcafe, &param->family_size, param->num_threads, globals.num_random_samples);
    }
    cafe_log(param, "Running Family-wide P-values ...\n");
    cafe_family_pvalues(param, param->max_pvalues);
    std::ofstream ofst((tokens[1] + ".pvalues").c_str());
    if (!ofst) throw std::runtime_error("ERROR(report): Cannot open " + tokens[1] + ".pvalues in write mode.\n");
    ofst << "ID\tFamily-wide P-value\n";
    for (size_t i = 0; i < param->pfamily->flist.size(); ++i) ofst << param->pfamily->flist[i].id << "\t" << param->max_pvalues[i] << "\n";
    viterbi_parameters viterbi;
    cafe_viterbi(param, viterbi);
    param->cutPvalues.clear();
    if (branchcutting) cafe_branch_cutting(param, globals.num_random_samples);  // cafe_do_report, reports.cpp:679-682
    if (likelihood) cafe_likelihood_ratio_test(param, param->max_pvalues.data());
    cafe_log(param, "Building Text report: %s\n", tokens[1].c_str());
    std::ofstream report((tokens[1] + ".cafe").c_str());
    if (!report) throw std::runtime_error("ERROR(report) : Cannot open " + tokens[1] + " in write mode.\n");
    cafe_report_text(report, param, viterbi);
    cafe_log(param, "Report Done\n");
    return 0;
}

int cafe_cmd_source(Globals& globals, std::vector<std::string> tokens) {
    if (tokens.size() != 2) throw std::runtime_error("Usage: source <file>\n");
    std::ifstream fp(tokens[1]);
    if (!fp) throw std::runtime_error("ERROR(source): Cannot open " + tokens[1] + " in read mode.\n");
    std::string line;
    int rtn = 0;
    while (std::getline(fp, line))
        if ((rtn = cafe_shell_dispatch_command(globals, line.c_str()))) break;
    return rtn;
}

static int cafe_cmd_date(Globals& globals, std::vector<std::string>) {
    time_t now = time(nullptr);
    cafe_log(&globals.param, "%s", ctime(&now));
    return 0;
}

std::map<std::string, cafe_command2> get_dispatcher() {
    std::map<std::string, cafe_command2> d;
    d["seed"] = cafe_cmd_seed;
    d["load"] = cafe_cmd_load;
    d["tree"] = cafe_cmd_tree;
    d["lambda"] = cafe_cmd_lambda;
    d["lambdamu"] = cafe_cmd_lambdamu;
    d["errormodel"] = cafe_cmd_errormodel;
    d["pvalue"] = cafe_cmd_pvalue;
    d["report"] = cafe_cmd_report;
    d["source"] = cafe_cmd_source;
    d["date"] = cafe_cmd_date;
    return d;
}

int cafe_shell_dispatch_command(Globals& globals, const char* cmd) {
    std::vector<std::string> tokens = tokenize(cmd);
    if (tokens.empty() || tokens[0][0] == '#') return 0;
    try {
        auto d = get_dispatcher();
        auto it = d.find(tokens[0]);
        if (it == d.end()) {
            std::fprintf(stderr, "cafe: %s: command not found (only the likelihood-path commands exist here)\n", tokens[0].c_str());
            return 0;
        }
        return it->second(globals, tokens);
    } catch (std::exception& ex) {
        std::fprintf(stderr, "%s\n", ex.what());
        return -1;
    }
}
