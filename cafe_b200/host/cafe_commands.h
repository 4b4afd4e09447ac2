// cafe_commands.h — the command entry points that reach the likelihood path, with the reference's
// signature `int cafe_cmd_X(Globals&, std::vector<std::string> tokens)` (cafe/cafe_commands.h:27) and
// argument meaning: seed, load, tree, lambda, lambdamu, errormodel, pvalue, report (text format, optional likelihood-ratio
// test), source.  Everything else of the reference's shell is out of scope.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "cafe_param.h"

struct Globals {  // cafe/Globals.h:13-33
    CafeParam param;
    int num_random_samples = 1000;
    Globals();
    ~Globals();
    void Clear(int btree_skip);
    void Prepare();  // cafe/Globals.cpp:153-171
};

struct Argument {  // cafe/cafe_commands.h
    std::string opt;
    std::vector<std::string> argv;
    int argc() const { return (int)argv.size(); }
};
std::vector<Argument> build_argument_list(std::vector<std::string> tokens);  // cafe_commands.cpp:476-501
std::vector<std::string> tokenize(const std::string& s);

typedef int (*cafe_command2)(Globals& globals, std::vector<std::string> tokens);
std::map<std::string, cafe_command2> get_dispatcher();

int cafe_cmd_seed(Globals& globals, std::vector<std::string> tokens);        // cafe_commands.cpp:1950
int cafe_cmd_load(Globals& globals, std::vector<std::string> tokens);        // cafe_commands.cpp:902
int cafe_cmd_tree(Globals& globals, std::vector<std::string> tokens);        // cafe_commands.cpp:1127
int cafe_cmd_lambda(Globals& globals, std::vector<std::string> tokens);      // lambda.cpp:369
int cafe_cmd_lambdamu(Globals& globals, std::vector<std::string> tokens);    // lambdamu.cpp:218
int cafe_cmd_errormodel(Globals& globals, std::vector<std::string> tokens);  // cafe_commands.cpp:1608
int cafe_cmd_pvalue(Globals& globals, std::vector<std::string> tokens);      // cafe_commands.cpp:1365
int cafe_cmd_report(Globals& globals, std::vector<std::string> tokens);      // cafe_commands.cpp:1010, reports.cpp:650-708
int cafe_cmd_source(Globals& globals, std::vector<std::string> tokens);      // cafe_commands.cpp:367
int cafe_shell_dispatch_command(Globals& globals, const char* cmd);          // cafe_commands.cpp:504
